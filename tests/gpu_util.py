"""Shared helpers of the -m gpu parity tests (all calls go through the C ABI via ctypes)."""
import numpy as np


def engine():
	from image_segmenter_b200.engine import get_engine

	return get_engine(0)


def to_dev(a, dtype=None):
	import torch

	t = torch.from_numpy(np.ascontiguousarray(a))
	if dtype is not None:
		t = t.to(dtype)
	return t.to(engine().dev)


def planes_of(X32):
	"""(n,3) float32 -> (3, npad) planar device tensor, 16-byte aligned rows."""
	import torch

	n = X32.shape[0]
	npad = (n + 3) & ~3
	p = torch.zeros((3, max(npad, 4)), dtype=torch.float32, device=engine().dev)
	if n:
		p[:, :n] = torch.from_numpy(np.ascontiguousarray(X32.T)).to(engine().dev)
	return p


def lloyd_step(planes, n, centers, *, exact=True, labels=True, inertia=False, fused=False, x2max=None, box=None):
	"""One cs_lloyd_step_f32 / cs_lloyd_iter_f32 call -> dict of host arrays.  box=None: the full walk over
	all centres; a box ((lo),(hi)) enables the grid-filtered assignment where the launch is eligible."""
	import torch
	from image_segmenter_b200 import _ffi

	e = engine()
	e.set_feature_box(box, 1 if box is not None else 0)  # a box in a test means: take the grid path from K = 4
	K = centers.shape[0]
	d_c = to_dev(np.asarray(centers, dtype=np.float64))
	d_lab = torch.full(((n + 3) & ~3,), 77, dtype=torch.uint8, device=e.dev) if labels else None
	d_sums = torch.zeros((K, 3), dtype=torch.float64, device=e.dev)
	d_cnt = torch.zeros(K, dtype=torch.float64, device=e.dev)
	flags = _ffi.CS_LLOYD_EXACT_TIES if exact else 0
	x2 = _ffi.CS_LAB_NORM2_MAX if x2max is None else float(x2max)
	lp = d_lab.data_ptr() if labels else None
	out = {}
	if fused:
		d_out = torch.zeros((K, 3), dtype=torch.float64, device=e.dev)
		d_stats = torch.zeros(4, dtype=torch.float64, device=e.dev)
		e._call("cs_lloyd_iter_f32", planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), n, d_c.data_ptr(), K,
		        lp, d_sums.data_ptr(), d_cnt.data_ptr(), d_out.data_ptr(), d_stats.data_ptr(), x2, flags)
		out["centers_new"], out["stats"] = d_out.cpu().numpy(), d_stats.cpu().numpy()
	else:
		d_in = torch.zeros(1, dtype=torch.float64, device=e.dev) if inertia else None
		e._call("cs_lloyd_step_f32", planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), n, d_c.data_ptr(), K,
		        lp, d_sums.data_ptr(), d_cnt.data_ptr(), d_in.data_ptr() if inertia else None, x2, flags)
		if inertia:
			out["inertia"] = float(d_in.item())
	torch.cuda.synchronize()
	if labels:
		full = d_lab.cpu().numpy()
		out["labels"], out["guard"] = full[:n], full[n:]
	out["sums"], out["counts"] = d_sums.cpu().numpy(), d_cnt.cpu().numpy()
	return out


def lab_like(rng, n):
	"""fp32 points spread like CIELAB values."""
	return np.stack([rng.uniform(0, 100, n), rng.uniform(-86, 98, n), rng.uniform(-107, 94, n)], 1).astype(np.float32)


def blobby_rgba(seed, h, w, ncol=6, sigma=9.0, alpha_holes=True, dark_corner=True):
	rng = np.random.default_rng(seed)
	cent = rng.integers(30, 256, (ncol, 3))
	which = (np.add.outer(np.arange(h) // max(h // 6, 1), np.arange(w) // max(w // 5, 1)) + rng.integers(0, 2, (h, w))) % ncol
	rgb = np.clip(cent[which] + rng.normal(0, sigma, (h, w, 3)), 0, 255).astype(np.uint8)
	a = np.full((h, w), 255, np.uint8)
	if alpha_holes:
		a[: h // 8] = 0
		a[h // 8: h // 6] = 100
		a[h // 6: h // 5, : w // 2] = 200
	if dark_corner:
		rgb[-h // 4:, -w // 4:] = rng.integers(0, 12, (len(rgb[-h // 4:]), len(rgb[0, -w // 4:]), 3))
	return np.dstack([rgb, a])


def kmeans_replay(X, K, n_init=10, centred=False, max_iter=300):
	"""KMeans(K, random_state=42, n_init=n_init).fit(X) replayed with the oracle's Lloyd from the
	product's k-means++ seeds.  centred=True mimics sklearn (Lloyd on X - mean: exact distance ties
	are then decided by rounding noise); centred=False is the documented tie rule (exact arithmetic,
	lowest index).  Returns (labels, float centres, inertia)."""
	from image_segmenter_b200 import color_simplify as cs
	from oracle import kmeans as okm

	X = np.ascontiguousarray(X, dtype=np.float64)
	mean = X.mean(axis=0) if centred else np.zeros(X.shape[1])
	Xc = X - mean
	tol = okm.sklearn_tol(X)
	best = None
	for idx in cs._seed_kmeans_plusplus(X, K, n_init):
		lab, inertia, cen, _ = okm.kmeans_single_lloyd(Xc, Xc[idx], max_iter, tol)
		same = best is not None and all(len(np.unique(best[0][lab == j])) <= 1 for j in range(K))
		if best is None or (inertia < best[2] and not same):
			best = (lab, cen + mean, inertia)
	return best
