"""Integer paths (K5-K8 + selection plumbing): bit-exact against NumPy / the oracle / Pillow."""
import ctypes as C

import numpy as np
import pytest

from oracle import mediancut as omc
from oracle import pipeline as op

from gpu_util import blobby_rgba, engine, to_dev

pytestmark = pytest.mark.gpu


def rgbkey(px):
	return (px[:, 0].astype(np.uint32) << 16) | (px[:, 1].astype(np.uint32) << 8) | px[:, 2].astype(np.uint32)


def rand_px(seed, n, low_entropy=False):
	rng = np.random.default_rng(seed)
	if low_entropy:
		pal = rng.integers(0, 256, (9, 4), dtype=np.uint8)
		return pal[rng.integers(0, 9, n)]
	return rng.integers(0, 256, (n, 4), dtype=np.uint8)


@pytest.mark.parametrize("n,low", [(1, False), (31, False), (33, True), (100003, False), (1 << 20, True), (3000001, False)])
def test_histogram_fold_compact(n, low):
	import torch

	e = engine()
	px = rand_px(n, n, low)
	d = to_dev(px)
	hist = torch.zeros(1 << 24, dtype=torch.int32, device=e.dev)
	e._call("cs_hist_rgb24", d.data_ptr(), n, hist.data_ptr())
	ref = np.bincount(rgbkey(px), minlength=1 << 24)
	assert np.array_equal(hist.cpu().numpy(), ref)
	for shift in (0, 2, 3):
		bits = 8 - shift
		nb = 1 << (3 * bits)
		cells = torch.empty(nb, dtype=torch.int32, device=e.dev)
		ncell = torch.zeros(1, dtype=torch.int32, device=e.dev)
		e._call("cs_hist_fold", hist.data_ptr(), shift, cells.data_ptr(), ncell.data_ptr())
		ck = ((px[:, 0].astype(np.uint32) >> shift) << (2 * bits)) | ((px[:, 1].astype(np.uint32) >> shift) << bits) | (px[:, 2].astype(np.uint32) >> shift)
		rc = np.bincount(ck, minlength=nb)
		assert np.array_equal(cells.cpu().numpy(), rc)
		nc = int(ncell.item())
		assert nc == int((rc > 0).sum())
		keys = torch.empty(nc, dtype=torch.int32, device=e.dev)
		counts = torch.empty(nc, dtype=torch.int32, device=e.dev)
		e._call("cs_hist_compact", cells.data_ptr(), nb, keys.data_ptr(), counts.data_ptr(), nc, ncell.data_ptr())
		assert int(ncell.item()) == nc
		assert np.array_equal(keys.cpu().numpy(), np.nonzero(rc)[0])
		assert np.array_equal(counts.cpu().numpy(), rc[rc > 0])


@pytest.mark.parametrize("seed,k", [(0, 2), (1, 5), (2, 16), (3, 64), (4, 256), (5, 100)])
@pytest.mark.parametrize("kind", ["blobby", "uniform", "few"])
def test_median_cut_bit_exact_vs_oracle_and_pillow(seed, k, kind):
	from PIL import Image

	e = engine()
	if kind == "blobby":
		img = blobby_rgba(seed, 150, 130)
	elif kind == "uniform":
		img = rand_px(seed, 320 * 300).reshape(300, 320, 4)  # > 65536 colours: cell shift >= 1
	else:
		img = rand_px(seed, 90 * 90, low_entropy=True).reshape(90, 90, 4)
	out, pal, idx = e.median_cut(to_dev(img.reshape(-1, 4)), k, True)
	pal_ref, idx_ref = omc.quantize(np.ascontiguousarray(img[:, :, :3]), k)
	assert np.array_equal(pal, pal_ref)
	assert np.array_equal(idx.cpu().numpy().reshape(img.shape[:2]), idx_ref)
	im = Image.fromarray(np.ascontiguousarray(img[:, :, :3])).quantize(colors=k, method=Image.Quantize.MEDIANCUT)
	assert np.array_equal(np.array(im.convert("RGB")), out.cpu().numpy().reshape(img.shape)[:, :, :3])
	assert np.array_equal(out.cpu().numpy().reshape(img.shape)[:, :, 3], img[:, :, 3])


@pytest.mark.parametrize("step", [1, 36, 85, 128, 255, 256])
@pytest.mark.parametrize("preserve_alpha", [True, False])
def test_posterize(step, preserve_alpha):
	e = engine()
	img = blobby_rgba(3, 77, 91)
	out, pal = e.posterize(to_dev(img.reshape(-1, 4)), step, preserve_alpha)
	q = (img[:, :, :3] // step) * step if step < 256 else np.zeros_like(img[:, :, :3])
	a = img[:, :, 3] if preserve_alpha else (img[:, :, 3] > 128).astype(np.uint8) * 255
	assert np.array_equal(out.cpu().numpy().reshape(img.shape), np.dstack([q, a]))
	assert np.array_equal(pal, np.unique(q.reshape(-1, 3), axis=0))


@pytest.mark.parametrize("kind", ["blobby", "uniform", "transparent"])
def test_statistics(kind):
	e = engine()
	img = blobby_rgba(4, 120, 100) if kind != "uniform" else rand_px(9, 256 * 256).reshape(256, 256, 4)
	if kind == "transparent":
		img[..., 3] = 0
	n_unique, n_op, s1, s2 = e.statistics(to_dev(img.reshape(-1, 4)))
	ref = op.statistics(img)
	assert n_unique == ref["total_unique_colors"] and n_op == ref["non_transparent_pixels"]
	sel = img[img[..., 3] > 0][:, :3].astype(np.int64)
	assert s1 == list(sel.sum(0)) and s2 == list((sel * sel).sum(0))


def test_mask_stats_and_unique():
	e = engine()
	img = blobby_rgba(8, 140, 90)
	d = to_dev(img.reshape(-1, 4))
	px = img.reshape(-1, 4)
	s = px[:, :3].astype(int).sum(1)
	op_ = px[:, 3] > 0
	for thr in (-1, 30, 90):
		n_op, n_hi, n_lo, nu = e.mask_stats(d, thr, want_unique=True)
		assert (n_op, n_hi, n_lo) == (int(op_.sum()), int((op_ & (s > 90)).sum()), int((op_ & (s > 30)).sum()))
		assert nu == len(np.unique(px[op_ & (s > thr)][:, :3], axis=0))
	hsva = e.rgba_to_hsv(d)
	h = hsva.cpu().numpy()
	for thr in (-1, 10, 30):
		n_op, n_hi, n_lo, nu = e.mask_stats(hsva, thr, hsv=True, want_unique=True)
		v = h[:, 2].astype(int)
		assert (n_op, n_hi, n_lo) == (int(op_.sum()), int((op_ & (v > 30)).sum()), int((op_ & (v > 10)).sum()))
		assert nu == len(np.unique(h[op_ & (v > thr)][:, :3], axis=0))


@pytest.mark.parametrize("n", [1, 255, 4096, 4097, 300007])
def test_select_compact_gather_channel_hist(n):
	e = engine()
	px = rand_px(n + 5, n)
	px[::3, 3] = 0
	d = to_dev(px)
	s = px[:, :3].astype(int).sum(1)
	for mode, thr in ((0, -1), (0, 300), (1, 128)):
		keep = (px[:, 3] > 0) & ((s > thr) if mode == 0 else (px[:, 2].astype(int) > thr))
		assert e.select_count(d, mode, thr) == int(keep.sum())
		out, idx = e.select_compact(d, mode, thr, want_index=True)
		assert np.array_equal(out.cpu().numpy(), px[keep])
		assert np.array_equal(idx.cpu().numpy(), np.nonzero(keep)[0])
		hist = e.channel_hist(d, mode, thr)
		ref = np.stack([np.bincount(px[keep][:, c], minlength=256) for c in range(3)])
		assert np.array_equal(hist, ref)
	ind = np.random.default_rng(0).integers(0, n, 1000)
	assert np.array_equal(e.gather(d, ind).cpu().numpy(), px[ind])


def test_sum_by_label_merge_remap_cooccurrence():
	e = engine()
	rng = np.random.default_rng(12)
	n, K = 50001, 256  # K = 256: label 255 is a real cluster, validity must come from the selection
	px = rand_px(1, n)
	px[rng.random(n) < 0.2, 3] = 0
	lab = rng.integers(0, K, n).astype(np.uint8)
	d, dl = to_dev(px), to_dev(lab)
	sel = (d, 0, 200)
	keep = (px[:, 3] > 0) & (px[:, :3].astype(int).sum(1) > 200)
	acc = e.sum_by_label(d, dl, K, sel=sel)
	ref = np.zeros((K, 4), np.int64)
	np.add.at(ref, lab[keep], np.concatenate([px[keep][:, :3].astype(np.int64), np.ones((keep.sum(), 1), np.int64)], 1))
	assert np.array_equal(acc, ref)
	fb = rng.integers(0, K, n).astype(np.uint8)
	merged = e.merge_labels(dl, to_dev(fb), n, sel=sel).cpu().numpy()
	assert np.array_equal(merged, np.where(keep, lab, fb))
	pal = rng.integers(0, 256, (K, 3), dtype=np.uint8)
	out = e.remap_labels(d, dl, pal, False, sel=sel).cpu().numpy()
	exp = np.zeros((n, 4), np.uint8)
	exp[keep, :3] = pal[lab[keep]]
	exp[:, 3] = (px[:, 3] > 128) * 255
	assert np.array_equal(out, exp)
	# sentinel convention (no selection pixels): label 255 = masked when K < 256
	lab2 = np.where(keep, lab % 7, 255).astype(np.uint8)
	out2 = e.remap_labels(d, to_dev(lab2), pal[:7], True).cpu().numpy()
	exp2 = np.zeros((n, 4), np.uint8)
	exp2[keep, :3] = pal[:7][lab2[keep]]
	exp2[:, 3] = px[:, 3]
	assert np.array_equal(out2, exp2)
	# same clustering: a permutation of the labels is the same clustering, a merge is not symmetric
	perm = rng.permutation(K).astype(np.uint8)
	assert e.same_clustering(dl, to_dev(perm[lab]), n, sel=sel)
	assert e.same_clustering(dl, to_dev((lab // 2).astype(np.uint8)), n, sel=sel)  # every label of l1 maps to one of l2
	assert not e.same_clustering(to_dev((lab // 2).astype(np.uint8)), dl, n, sel=sel)


def test_full_size_16mp_histogram_properties():
	"""BASELINE config 5 size (4096 x 4096): checksum properties without a CPU pass over the output —
	histogram total = n, fold preserves the total, box sums add up to the channel totals."""
	import torch

	e = engine()
	n = 4096 * 4096
	g = torch.Generator(device=e.dev)
	g.manual_seed(5)
	d = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=e.dev, generator=g)
	hist = torch.zeros(1 << 24, dtype=torch.int32, device=e.dev)
	e._call("cs_hist_rgb24", d.data_ptr(), n, hist.data_ptr())
	assert int(hist.sum(dtype=torch.int64).item()) == n
	out, pal, idx = e.median_cut(d, 256, True)
	assert len(pal) == 256 and int(idx.max().item()) == 255
	assert torch.equal(out[:, 3], d[:, 3])
	# every output colour is a palette colour and the map is idempotent
	pal_t = torch.from_numpy(pal).to(e.dev)
	assert torch.equal(out[:, :3], pal_t[idx.long()])
	sub = out[::1024].cpu().numpy()
	src = d[::1024].cpu().numpy()
	dist = ((src[:, None, :3].astype(int) - pal[None].astype(int)) ** 2).sum(-1)
	chosen = ((src[:, :3].astype(int) - sub[:, :3].astype(int)) ** 2).sum(-1)
	assert np.array_equal(chosen, dist.min(1))  # the chosen entry attains the minimum distance
