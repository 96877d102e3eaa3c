"""K2/K3 parity: the CUDA Lloyd step (through the C ABI) against the oracle on the same seeded
inputs and initial centres.  Labels: identical in EXACT_TIES mode (fp64 first-minimum); in the
fast mode a label may differ only where the two best distances are within the documented fp32
bound.  Sums / counts: consistent with the labels the kernel chose, centres within 1e-4 relative
(north star) — in fact ~1e-7."""
import numpy as np
import pytest

from oracle import kmeans as okm

from gpu_util import engine, lab_like, lloyd_step, planes_of, to_dev

pytestmark = pytest.mark.gpu

REL = 1e-4  # north-star tolerance for centroids / LAB values


def _direct_gap(X, C):
	best, second = okm.near_tie_gap(X.astype(np.float64), C)
	return second - best


def _check_step(X32, C, exact, fused=False, box=None):
	n, K = len(X32), len(C)
	X = X32.astype(np.float64)
	r = lloyd_step(planes_of(X32), n, C, exact=exact, inertia=not fused, fused=fused, box=box)
	ref_lab = okm.assign_labels(X, C) if n else np.zeros(0, np.int32)
	lab = r["labels"]
	assert (r["guard"] == 77).all(), "kernel wrote past the label array"
	mism = np.nonzero(lab != ref_lab)[0]
	if len(mism):
		best, second = okm.near_tie_gap(X[mism], C)
		gap = second - best
		cn = (C ** 2).sum(1).max()
		if exact:
			bound = 1e-9
		elif K <= 64:  # integer keys: 7 roundings of magnitude <= (sqrt(x2max) + max|c|)^2 per key, two keys
			bound = 1.5 * 2 * 7 * 2.0 ** -24 * ((np.sqrt(31400.0) + np.sqrt(cn)) ** 2 + 1024)
		else:  # float keys: index bits replace the low mantissa bits of the distance
			bits = int(np.ceil(np.log2(K)))
			bound = 1.5 * 2.2 * (5 * 2.0 ** -24 * ((X[mism] ** 2).sum(1) + 2 * cn) + 2.0 ** -(23 - bits) * second)
		assert (gap <= bound).all(), f"{(gap > bound).sum()} label mismatches beyond the near-tie bound"
	# sums and counts follow the labels the kernel chose
	s2, c2 = okm.accumulate(X, lab.astype(np.int64), K) if n else (np.zeros((K, 3)), np.zeros(K))
	assert np.array_equal(r["counts"], c2)
	assert np.allclose(r["sums"], s2, rtol=2e-6, atol=1e-3)
	if not fused:
		ref_in = okm.inertia(X, C, lab.astype(np.int64)) if n else 0.0
		# fast mode recovers the distance from the fp32 key: absolute error per pixel bounded by the key
		# rounding (7 roundings of magnitude (sqrt(x2max) + max|c|)^2 ~ 1e5 -> ~0.05); exact mode recomputes it
		slack = 0.0 if exact else n * 7 * 2.0 ** -24 * (np.sqrt(31400.0) + np.sqrt((C ** 2).sum(1).max())) ** 2
		assert abs(r["inertia"] - ref_in) <= REL * max(ref_in, 1.0) + slack
	else:
		nz = c2 > 0
		exp = r["sums"][nz] * (1.0 / r["counts"][nz])[:, None]
		assert np.array_equal(r["centers_new"][nz], exp)
		assert r["stats"][1] == (K - nz.sum())
		if nz.all():
			assert abs(r["stats"][0] - okm.center_shift_total(C, r["centers_new"])) <= 1e-9 * max(1.0, r["stats"][0])
	return len(mism)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 2047, 2048, 2049, 100003])
@pytest.mark.parametrize("K", [2, 5, 8, 16, 17, 64, 200, 256])
def test_step_small(n, K):
	rng = np.random.default_rng(n * 1000 + K)
	X32 = lab_like(rng, n)
	C = X32[rng.choice(n, K, replace=(n < K))].astype(np.float64)
	if n < K:
		C = C + rng.normal(0, 1e-3, C.shape)
	_check_step(X32, C, exact=True)
	_check_step(X32, C, exact=False)
	_check_step(X32, C, exact=True, fused=True)


def test_empty_input():
	r = lloyd_step(planes_of(np.zeros((0, 3), np.float32)), 0, np.array([[1.0, 2, 3], [4, 5, 6]]), exact=True)
	assert (r["counts"] == 0).all() and (r["sums"] == 0).all()


def test_exact_ties_lowest_index_wins():
	"""Duplicate centres and points exactly between two centres: first minimum, as
	_k_means_lloyd.pyx:205-213."""
	C = np.array([[10.0, 0, 0], [20.0, 0, 0], [10.0, 0, 0], [30.0, 5, 5]])
	X32 = np.array([[10, 0, 0], [15, 0, 0], [20, 0, 0], [25, 2.5, 2.5]] * 300, dtype=np.float32)
	r = lloyd_step(planes_of(X32), len(X32), C, exact=True)
	ref = okm.assign_labels(X32.astype(np.float64), C)
	direct = np.array([0, 0, 1, 1] * 300)
	assert np.array_equal(r["labels"], direct)
	# the oracle's GEMM form may round the exact tie either way; where it is decisive both agree
	decided = _direct_gap(X32, C) > 1e-9
	assert np.array_equal(r["labels"][decided], ref[decided])


def test_golden_sklearn_step(golden_lloyd):
	g = golden_lloyd
	X32 = g["lab32"]
	r = lloyd_step(planes_of(X32), len(X32), g["C0"], exact=True, fused=True)
	assert np.array_equal(r["labels"], g["step_labels"].astype(np.uint8))
	assert np.array_equal(r["counts"], g["step_weights"])
	assert np.allclose(r["centers_new"], g["step_centers"], rtol=1e-6, atol=1e-6)
	assert abs(r["stats"][0] - float((g["step_shift"] ** 2).sum())) <= 1e-5 * max(1.0, r["stats"][0])


def test_golden_sklearn_full_fit(golden_lloyd):
	from image_segmenter_b200.engine import KMeansGPU

	g = golden_lloyd
	X32 = g["lab32"]
	e = engine()
	km = KMeansGPU(e, "f32", len(X32), planes=planes_of(X32))
	fit = km.fit_single(g["C0"], max_iter=300, tol=okm.sklearn_tol(X32.astype(np.float64)))
	assert fit.n_iter == int(g["fit_n_iter"])
	assert np.array_equal(fit.labels[:len(X32)].cpu().numpy(), g["fit_labels"].astype(np.uint8))
	assert np.allclose(fit.centers, g["fit_centers"], rtol=REL, atol=1e-5)
	assert abs(fit.inertia - float(g["fit_inertia"])) <= REL * float(g["fit_inertia"])


@pytest.mark.parametrize("K", [8, 16, 64])
def test_multi_iteration_trajectory_matches_oracle(K):
	"""20 iterations from the same C0: centres track the fp64 oracle within 1e-4 relative."""
	from image_segmenter_b200.engine import KMeansGPU

	rng = np.random.default_rng(K)
	n = 300000
	cent = lab_like(rng, K)
	X32 = (cent[rng.integers(0, K, n)] + rng.normal(0, 6, (n, 3))).astype(np.float32)
	C0 = X32[rng.choice(n, K, replace=False)].astype(np.float64)
	km = KMeansGPU(engine(), "f32", n, planes=planes_of(X32), x2max=float((X32.astype(np.float64) ** 2).sum(1).max()) * 1.01)
	fit = km.fit_single(C0, max_iter=20, tol=0.0)
	labels, inertia, centers, n_iter = okm.kmeans_single_lloyd(X32.astype(np.float64), C0, max_iter=20, tol=0.0)
	assert fit.n_iter == n_iter
	assert np.allclose(fit.centers, centers, rtol=REL, atol=REL)
	lab = fit.labels[:n].cpu().numpy()
	assert (lab != labels.astype(np.uint8)).sum() <= 2  # documented: fp64 ties of the GEMM form only
	assert abs(fit.inertia - inertia) <= REL * inertia


def test_relocation_of_empty_cluster():
	rng = np.random.default_rng(9)
	X32 = lab_like(rng, 5000)
	C = np.concatenate([X32[:3].astype(np.float64), [[1e4, 1e4, 1e4]]])  # the last centre attracts nothing
	e = engine()
	import torch

	P = planes_of(X32)
	r = lloyd_step(P, len(X32), C, exact=True, x2max=4e8)
	assert r["counts"][3] == 0
	d_lab, d_c = to_dev(np.pad(r["labels"], (0, 3))), to_dev(C)
	d_s, d_n = to_dev(r["sums"]), to_dev(r["counts"])
	e._call("cs_lloyd_relocate_f32", P[0].data_ptr(), P[1].data_ptr(), P[2].data_ptr(), len(X32), d_lab.data_ptr(),
	        d_c.data_ptr(), 4, d_s.data_ptr(), d_n.data_ptr())
	torch.cuda.synchronize()
	s_ref, c_ref = okm.relocate_empty(X32.astype(np.float64), C, r["sums"], r["counts"], r["labels"].astype(np.int64))
	assert np.array_equal(d_n.cpu().numpy(), c_ref)
	assert np.allclose(d_s.cpu().numpy(), s_ref, rtol=1e-12, atol=1e-9)


def test_rgba8_step_exact_integer_sums():
	rng = np.random.default_rng(11)
	n, K = 200001, 12
	px = rng.integers(0, 256, (n, 4), dtype=np.uint8)
	px[rng.random(n) < 0.1, 3] = 0
	C = px[rng.choice(n, K, replace=False), :3].astype(np.float64) + 0.25
	e = engine()
	import torch

	d = to_dev(px)
	for thr in (-1, 90):
		d_lab = torch.full(((n + 3) & ~3,), 77, dtype=torch.uint8, device=e.dev)
		d_s, d_n = torch.zeros((K, 3), dtype=torch.float64, device=e.dev), torch.zeros(K, dtype=torch.float64, device=e.dev)
		d_i = torch.zeros(1, dtype=torch.float64, device=e.dev)
		e._call("cs_lloyd_step_rgba8", d.data_ptr(), n, thr, to_dev(C).data_ptr(), K, d_lab.data_ptr(), d_s.data_ptr(),
		        d_n.data_ptr(), d_i.data_ptr(), 1)
		torch.cuda.synchronize()
		keep = (px[:, 3] > 0) & (px[:, :3].astype(int).sum(1) > thr)
		X = px[:, :3].astype(np.float64)
		ref = okm.assign_labels(X[keep], C)
		lab = d_lab.cpu().numpy()[:n]
		assert (lab[~keep] == 255).all()
		mism = np.nonzero(lab[keep] != ref)[0]
		if len(mism):  # only exact fp64 ties of the GEMM form
			assert (_direct_gap(X[keep][mism], C) <= 1e-9).all()
		s2, c2 = okm.accumulate(X[keep], lab[keep].astype(np.int64), K)
		assert np.array_equal(d_s.cpu().numpy(), s2) and np.array_equal(d_n.cpu().numpy(), c2)  # exact integers
		ref_in = okm.inertia(X[keep], C, lab[keep].astype(np.int64))
		assert abs(float(d_i.item()) - ref_in) <= 1e-5 * ref_in


@pytest.mark.parametrize("K", [16, 64])
def test_full_size_properties_64mp(K):
	"""BASELINE config 3 size (8192 x 8192): properties that need no CPU pass over 64 MP —
	counts sum to n, per-plane sums add up to the plane totals (linearity), labels < K, and
	re-assigning with the same centres is idempotent; a strided 1/64 sample is checked against
	the oracle."""
	import torch

	e = engine()
	n = 8192 * 8192
	g = torch.Generator(device=e.dev)
	g.manual_seed(3)
	P = torch.empty((3, n), dtype=torch.float32, device=e.dev)
	for j, (s, o) in enumerate(((100, 0), (184, -86), (201, -107))):
		P[j] = torch.rand(n, device=e.dev, generator=g) * s + o
	idx = torch.randint(0, n, (K,), device=e.dev, generator=g)
	C = P[:, idx].T.double().cpu().numpy()
	r1 = lloyd_step(P, n, C, exact=True)
	assert r1["counts"].sum() == n
	assert r1["labels"].max() < K
	tot = P.double().sum(dim=1).cpu().numpy()
	assert np.allclose(r1["sums"].sum(0), tot, rtol=1e-6)
	r2 = lloyd_step(P, n, C, exact=True)
	assert np.array_equal(r1["labels"], r2["labels"]) and np.array_equal(r1["sums"], r2["sums"])  # deterministic
	sub = slice(0, n, 64)
	Xs = P[:, sub].T.cpu().numpy().astype(np.float64)
	ref = okm.assign_labels(Xs, C)
	mism = np.nonzero(r1["labels"][sub] != ref)[0]
	if len(mism):
		assert (_direct_gap(Xs[mism], C) <= 1e-9).all()
	rf = lloyd_step(P, n, C, exact=False, labels=False)
	assert rf["counts"].sum() == n
	assert np.allclose(rf["sums"] / np.maximum(rf["counts"], 1)[:, None], r1["sums"] / np.maximum(r1["counts"], 1)[:, None],
	                   rtol=REL, atol=REL)


@pytest.mark.parametrize("kind", ["f32", "rgba8"])
def test_fit_with_empty_cluster_midrun_matches_oracle(kind):
	"""The batched device-side loop control stops at the iteration that finds an empty cluster, the host
	relocates (farthest point, as _relocate_empty_clusters_dense) and the run continues — same trajectory
	as the oracle's loop."""
	from image_segmenter_b200.engine import KMeansGPU

	rng = np.random.default_rng(17)
	n, K = 6000, 6
	if kind == "f32":
		X32 = lab_like(rng, n)
		X = X32.astype(np.float64)
		km = KMeansGPU(engine(), "f32", n, planes=planes_of(X32))
	else:
		px = rng.integers(0, 256, (n, 4), dtype=np.uint8)
		px[:, 3] = 255
		X = px[:, :3].astype(np.float64)
		km = KMeansGPU(engine(), "rgba8", n, px=to_dev(px), mask_mode=0, min_bright=-1)
	C0 = X[rng.choice(n, K, replace=False)].copy()
	C0[4] = [900.0, 900.0, 900.0]  # attracts nothing in the first iteration
	if kind == "f32":
		km.x2max = 4.0e6
	fit = km.fit_single(C0, max_iter=40, tol=okm.sklearn_tol(X))
	labels, inertia, centers, n_iter = okm.kmeans_single_lloyd(X, C0, max_iter=40, tol=okm.sklearn_tol(X))
	assert fit.n_iter == n_iter
	assert np.allclose(fit.centers, centers, rtol=REL, atol=REL)
	assert (fit.labels[:n].cpu().numpy() != labels.astype(np.uint8)).sum() <= 2
	assert (fit.counts > 0).all()


@pytest.mark.parametrize("n,K", [(100003, 16), (8 * 1024 * 1024 + 5, 16), (300007, 64)])
def test_chained_launches_equal_unchained(n, K):
	"""CS_LLOYD_CHAINED (programmatic dependent launch: prologue and first tile loads under the previous
	launch's tail) must not change a single bit: 12 ping-pong iterations queued without host
	synchronisation, once with and once without the flag, from the same centres."""
	import torch
	from image_segmenter_b200 import _ffi

	e = engine()
	rng = np.random.default_rng(n + K)
	X32 = lab_like(rng, n)
	P = planes_of(X32)
	C0 = X32[rng.choice(n, K, replace=False)].astype(np.float64)
	res = []
	for chained in (0, _ffi.CS_LLOYD_CHAINED):
		c = [to_dev(C0.copy()), to_dev(np.zeros_like(C0))]
		d_lab = torch.zeros(((n + 3) & ~3,), dtype=torch.uint8, device=e.dev)
		d_sums = torch.zeros((K, 3), dtype=torch.float64, device=e.dev)
		d_cnt = torch.zeros(K, dtype=torch.float64, device=e.dev)
		d_stats = torch.zeros(4, dtype=torch.float64, device=e.dev)
		torch.cuda.synchronize()
		for it in range(12):
			e._call("cs_lloyd_iter_f32", P[0].data_ptr(), P[1].data_ptr(), P[2].data_ptr(), n, c[it & 1].data_ptr(), K,
			        d_lab.data_ptr(), d_sums.data_ptr(), d_cnt.data_ptr(), c[(it & 1) ^ 1].data_ptr(), d_stats.data_ptr(),
			        _ffi.CS_LAB_NORM2_MAX, chained if it else 0)
		torch.cuda.synchronize()
		res.append([t.cpu().numpy() for t in (c[0], c[1], d_lab, d_sums, d_cnt, d_stats)])
	for a, b in zip(*res):
		assert np.array_equal(a, b)
	assert res[0][4].sum() == n


def test_chained_batch_halts_and_drains():
	"""cs_lloyd_run_f32 chains its launches; once the device-side control block says "converged" the
	remaining launches of the batch return at once (after their prefetched tiles have landed) and leave
	centres, iteration count and statistics untouched."""
	import torch
	from image_segmenter_b200 import _ffi

	e = engine()
	rng = np.random.default_rng(11)
	n, K = 200001, 8
	# eight tight, well separated blobs: converges in a few iterations
	cen = lab_like(rng, K)
	X32 = (cen[rng.integers(0, K, n)] + rng.normal(0, 0.5, (n, 3))).astype(np.float32)
	P = planes_of(X32)
	C0 = X32[rng.choice(n, K, replace=False)].astype(np.float64)
	_, _, cref, it_ref = okm.kmeans_single_lloyd(X32.astype(np.float64), C0, max_iter=40, tol=1e-6)
	a, b = to_dev(C0.copy()), to_dev(np.zeros_like(C0))
	d_sums = torch.zeros((K, 3), dtype=torch.float64, device=e.dev)
	d_cnt = torch.zeros(K, dtype=torch.float64, device=e.dev)
	d_stats = torch.zeros(4, dtype=torch.float64, device=e.dev)
	ctl = torch.tensor([0.0, 0.0, 1e-6, 0.0], dtype=torch.float64, device=e.dev)
	e._call("cs_lloyd_run_f32", P[0].data_ptr(), P[1].data_ptr(), P[2].data_ptr(), n, a.data_ptr(), b.data_ptr(), K,
	        d_sums.data_ptr(), d_cnt.data_ptr(), d_stats.data_ptr(), _ffi.CS_LAB_NORM2_MAX, _ffi.CS_LLOYD_EXACT_TIES, 40,
	        ctl.data_ptr())
	torch.cuda.synchronize()
	h = ctl.cpu().numpy()
	assert h[0] == 1.0 and 1 <= h[1] < 40
	assert int(h[1]) == it_ref
	final = (b if int(h[1]) & 1 else a).cpu().numpy()
	assert np.allclose(final, cref, rtol=REL, atol=REL)
	assert d_cnt.sum().item() == n
