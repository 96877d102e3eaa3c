"""The drop-in module `image_segmenter_b200.color_simplify` against the fixtures made by the
UNMODIFIED reference (tests/golden/reference_entry_points.npz) and against the oracle on larger
seeded images.  Integer paths bit-exact; k-means palettes equal up to the documented +-1 LSB
truncation artefact (SURVEY §0.3, §8c iv)."""
import warnings

import numpy as np
import pytest

from oracle import pipeline as op

from gpu_util import blobby_rgba

pytestmark = pytest.mark.gpu

IMAGES = ("blobby", "uniform", "fewcolors")


@pytest.fixture(scope="module")
def cs():
	from image_segmenter_b200 import color_simplify

	return color_simplify


def _eq(g, tag, out, pal):
	assert np.array_equal(out, g[f"{tag}__rgba"]), tag
	ref = g[f"{tag}__palette"]
	assert np.array_equal(np.asarray(pal), ref) and np.asarray(pal).dtype == ref.dtype, tag


def _sse(X, pal):
	return float(((X[:, None, :] - np.asarray(pal, dtype=np.float64)[None]) ** 2).sum(-1).min(1).sum())


def _kmeans_palette_ok(img, pal, ref_pal, key):
	d = pal.astype(int) - ref_pal.astype(int)
	if ((d == 0) | (d == 1)).all():
		return
	from gpu_util import kmeans_replay

	px = img.reshape(-1, 4)
	s = px[:, :3].astype(int).sum(1)
	keep = (px[:, 3] > 0) & (s > 90)
	if keep.sum() < len(ref_pal):
		keep = (px[:, 3] > 0) & (s > 30)
	if keep.sum() == 0:
		keep = px[:, 3] > 0
	X = px[keep][:, :3].astype(np.float64)
	_, cen, _ = kmeans_replay(X, len(ref_pal), centred=False)
	rep = np.clip(cen, 0, 255).astype(np.uint8)
	dd = pal.astype(int) - rep.astype(int)
	assert ((dd == 0) | (dd == 1)).all() or abs(_sse(X, pal) - _sse(X, ref_pal)) <= 0.02 * max(_sse(X, ref_pal), 1.0), key


def _hsv_result_ok(img, out, pal, ref_out, ref_pal, key):
	if np.array_equal(out, ref_out) and np.array_equal(np.asarray(pal), ref_pal):
		return
	# a tie went the other way: same number of colours, every output pixel carries a palette colour, and the
	# palette explains the opaque pixels as well as the reference's does
	px = img.reshape(-1, 4)
	op_ = px[:, 3] > 0
	X = px[op_][:, :3].astype(np.float64)
	got = out.reshape(-1, 4)[op_][:, :3]
	assert {tuple(c) for c in got} <= {tuple(c) for c in np.asarray(pal)}, key
	assert abs(_sse(X, pal) - _sse(X, ref_pal)) <= 0.05 * max(_sse(X, ref_pal), 1.0), key


def _palette_close(pal, ref):
	"""uint8 palettes from truncated float centres: equal, or +1 where the reference's fp64 sum
	landed just below an exact integer (153.9999... -> 153)."""
	assert pal.shape == ref.shape and pal.dtype == ref.dtype
	d = pal.astype(int) - ref.astype(int)
	assert ((d == 0) | (d == 1)).all(), (pal, ref)


@pytest.mark.parametrize("name", IMAGES)
def test_integer_entry_points_bit_exact(golden, cs, name):
	img = golden[f"in_{name}"]
	before = img.copy()
	for k in (8, 16, 100):
		_eq(golden, f"{name}__median_cut_{k}", *cs.simplify_colors_median_cut(img, k))
	for k in (6, 16, 256):
		_eq(golden, f"{name}__octree_{k}", *cs.simplify_colors_octree(img, k))
	for k in (2, 8, 16, 256):
		_eq(golden, f"{name}__threshold_{k}", *cs.simplify_colors_threshold(img, k))
	_eq(golden, f"{name}__threshold_8_noalpha", *cs.simplify_colors_threshold(img, 8, preserve_alpha=False))
	assert np.array_equal(img, before)  # input never mutated


@pytest.mark.parametrize("name", IMAGES)
def test_statistics(golden, cs, name):
	st = cs.get_color_statistics(golden[f"in_{name}"])
	ref = golden[f"{name}__stats"]
	assert st["total_unique_colors"] == int(ref[0]) and isinstance(st["total_unique_colors"], int)
	assert st["non_transparent_pixels"] == int(ref[1]) and isinstance(st["non_transparent_pixels"], np.int64)
	assert np.allclose(st["rgb_mean"], ref[2:5], rtol=1e-12) and np.allclose(st["rgb_std"], ref[5:8], rtol=1e-10)
	assert st["image_size"] == golden[f"in_{name}"].shape[:2]


@pytest.mark.parametrize("name", IMAGES)
def test_custom_palette(golden, cs, name):
	img, cp = golden[f"in_{name}"], golden["custom_palette_in"]
	for metric in ("rgb", "hsv", "lab"):
		out, pal = cs.simplify_colors_custom_palette(img, cp, True, metric)
		assert pal is cp
		_eq(golden, f"{name}__custom_{metric}", out, pal)
	_eq(golden, f"{name}__custom_lab_noalpha", *cs.simplify_colors_custom_palette(img, cp, False, "lab"))


@pytest.mark.parametrize("name", IMAGES)
def test_kmeans_vs_reference(golden, cs, name):
	"""RGB k-means.  Integer pixels and integer k-means++ seeds make EXACT distance ties common; the
	kernel resolves them in exact arithmetic to the lowest index (as _k_means_lloyd.pyx:205-213 would
	on exact values), while sklearn's own result on such pixels depends on the rounding noise of its
	centred GEMM form (and on its thread count).  So: (a) always equal to the oracle replay with the
	documented tie rule; (b) equal to the reference's fixture whenever no tie was decided differently
	(checked by replaying both ways on the CPU); (c) otherwise of the same quality."""
	from gpu_util import kmeans_replay

	img = golden[f"in_{name}"]
	px = img.reshape(-1, 4)
	for k in (5, 16):
		out, pal = cs.simplify_colors_kmeans(img, k, strict_reference_quirks=True)
		ref_pal = golden[f"{name}__kmeans_{k}__palette"]
		assert np.array_equal(out, golden[f"{name}__kmeans_{k}__rgba"])  # the reference's all-zero RGB + alpha
		s = px[:, :3].astype(int).sum(1)
		keep = (px[:, 3] > 0) & (s > 90)
		assert keep.sum() >= k
		X = px[keep][:, :3].astype(np.float64)
		lab_u, cen_u, in_u = kmeans_replay(X, len(ref_pal), centred=False)
		_palette_close(pal, np.clip(cen_u, 0, 255).astype(np.uint8))
		lab_c, cen_c, in_c = kmeans_replay(X, len(ref_pal), centred=True)
		if np.array_equal(lab_u, lab_c):
			_palette_close(pal, ref_pal)
		else:
			in_ref = ((X[:, None, :] - ref_pal[None].astype(np.float64)) ** 2).sum(-1).min(1).sum()
			in_gpu = ((X[:, None, :] - pal[None].astype(np.float64)) ** 2).sum(-1).min(1).sum()
			assert abs(in_gpu - in_ref) <= 0.02 * in_ref
		# intended remap: every kept pixel carries its centre, everything else is RGB 0
		out2, pal2 = cs.simplify_colors_kmeans(img, k)
		assert np.array_equal(pal2, pal)
		exp = np.zeros_like(px)
		exp[keep, :3] = pal[lab_u]
		exp[:, 3] = px[:, 3]
		assert np.array_equal(out2.reshape(-1, 4), exp)
	out, pal = cs.simplify_colors_adaptive(img, 4, True, "kmeans")
	assert pal.shape == golden[f"{name}__adaptive_kmeans_4__palette"].shape


@pytest.mark.parametrize("name", IMAGES)
def test_hsv_clustering_vs_reference(golden, cs, name):
	img = golden[f"in_{name}"]
	out, pal = cs.simplify_colors_hsv_clustering(img, 6)
	ref_out, ref_pal = golden[f"{name}__hsv_6__rgba"], golden[f"{name}__hsv_6__palette"]
	assert np.array_equal(pal, ref_pal)  # centres are exact integer means: no truncation artefact
	assert np.array_equal(out, ref_out)


@pytest.mark.parametrize("name", IMAGES)
def test_perceptual_vs_reference(golden, cs, name):
	img = golden[f"in_{name}"]
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		np.random.seed(7)
		_eq(golden, f"{name}__perceptual_fast_6", *cs.simplify_colors_perceptual_fast(img, 6))
		np.random.seed(7)
		_eq(golden, f"{name}__perceptual_5", *cs.simplify_colors_perceptual(img, 5, max_samples=2000))


def test_degenerate_inputs_return_input_object(cs):
	img = np.zeros((6, 5, 4), np.uint8)  # fully transparent
	for fn in (cs.simplify_colors_kmeans, cs.simplify_colors_perceptual, cs.simplify_colors_perceptual_fast,
	           cs.simplify_colors_hsv_clustering, cs.simplify_colors_adaptive_distance):
		out, pal = fn(img)
		assert out is img and np.array_equal(pal, [[0, 0, 0]]) and pal.dtype == np.int64
	one = np.zeros((6, 5, 4), np.uint8)
	one[...] = (200, 100, 50, 255)  # a single colour: K collapses to 1 < 2
	for fn in (cs.simplify_colors_kmeans, cs.simplify_colors_hsv_clustering, cs.simplify_colors_perceptual):
		out, pal = fn(one)
		assert out is one and np.array_equal(pal, [[0, 0, 0]])
	st = cs.get_color_statistics(img)
	assert st["non_transparent_pixels"] == 0 and np.array_equal(st["rgb_mean"], [0, 0, 0])
	cp = np.array([[1, 2, 3]], np.uint8)
	out, pal = cs.simplify_colors_custom_palette(img, cp)
	assert out is img and pal is cp


def test_adaptive_dispatch(cs):
	img = blobby_rgba(21, 64, 64, ncol=3, sigma=0.0, dark_corner=False)
	# few unique colours (<= target): "adaptive" -> threshold
	out, pal = cs.simplify_colors_adaptive(img, 64, True, "adaptive")
	ref, rp = op.threshold(img, 64)
	assert np.array_equal(out, ref) and np.array_equal(pal, rp)
	# unknown algorithm id -> k-means (reference :340-342)
	out, pal = cs.simplify_colors_adaptive(img, 3, True, "no_such_algorithm")
	assert pal.shape == (3, 3)


@pytest.mark.parametrize("shape,k", [((517, 389), 16), ((1080, 1920), 8)])
def test_larger_images_vs_oracle(cs, shape, k):
	img = blobby_rgba(31, *shape)
	for fn, ofn in ((cs.simplify_colors_median_cut, lambda i, kk: op.median_cut(i, kk)),
	                (cs.simplify_colors_octree, lambda i, kk: op.median_cut(i, kk, power_of_two=False)),
	                (cs.simplify_colors_threshold, op.threshold)):
		out, pal = fn(img, k)
		ro, rp = ofn(img, k)
		assert np.array_equal(out, ro) and np.array_equal(pal, rp)
	cp = np.random.default_rng(1).integers(0, 256, (k, 3), dtype=np.uint8)
	for metric in ("rgb", "hsv"):
		out, _ = cs.simplify_colors_custom_palette(img, cp, True, metric)
		ro, _ = op.custom_palette(img, cp, True, metric)
		assert np.array_equal(out, ro)
	out, _ = cs.simplify_colors_custom_palette(img, cp, True, "lab")
	ro, _ = op.custom_palette(img, cp, True, "lab")
	assert (out != ro).any(axis=2).sum() <= 4  # fp64 rounding-level ties only
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		np.random.seed(3)
		out, pal = cs.simplify_colors_perceptual_fast(img, k)
		np.random.seed(3)
		ro, rp = op.perceptual_fast(img, k)
	assert np.array_equal(pal, rp)
	assert (out != ro).any(axis=2).sum() <= 4
	st, rs = cs.get_color_statistics(img), op.statistics(img)
	assert st["total_unique_colors"] == rs["total_unique_colors"]
	assert np.allclose(st["rgb_std"], rs["rgb_std"], rtol=1e-10)


def test_adaptive_distance_vs_reference(golden, cs):
	"""DBSCAN path (host scikit-learn, as the reference) around the device LAB conversion, per-label sums
	and gather: equal to the unmodified reference where it runs, IndexError where its merge branch breaks."""
	from pathlib import Path

	ga = np.load(Path(__file__).resolve().parent / "golden" / "reference_adaptive_distance.npz")
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		for name, k in (("blobby", 3), ("blobby", 6), ("blobby", 8), ("fewcolors", 8)):
			out, pal = cs.simplify_colors_adaptive_distance(golden[f"in_{name}"], k)
			assert np.array_equal(pal, ga[f"{name}__ad_{k}__palette"]) and pal.dtype == np.uint8, (name, k)
			assert np.array_equal(out, ga[f"{name}__ad_{k}__rgba"]), (name, k)
		for name, k in (("uniform", 6), ("fewcolors", 3)):
			assert f"{name}__ad_{k}__indexerror" in ga.files
			with pytest.raises(IndexError):
				cs.simplify_colors_adaptive_distance(golden[f"in_{name}"], k)


@pytest.mark.parametrize("name,k", [("blobby", 5), ("blobby", 16), ("uniform", 8), ("fewcolors", 7)])
def test_device_kmeanspp_matches_sklearn_stream(golden, cs, name, k):
	"""The device k-means++ passes draw the same seeds as sklearn's _kmeans_plusplus with RandomState(42),
	for all 10 initialisations, on RGB rows and on the weighted HSV feature rows."""
	from image_segmenter_b200 import _colorspace as csp
	from image_segmenter_b200.engine import get_engine

	eng = get_engine()
	img = golden[f"in_{name}"]
	d = eng.upload_rgba(img)
	px, _ = eng.select_compact(d, 0, 90)
	rows = px.cpu().numpy()
	X = rows[:, :3].astype(np.float64)
	ident = np.tile(np.arange(256, dtype=np.float64), (3, 1))
	idx, cents = eng.kmeanspp_seeds(px, ident, k, 10)
	ref = cs._seed_kmeans_plusplus(X, k, 10)
	for a, b, c in zip(idx, ref, cents):
		# the same rows; where a colour occurs many times sklearn may pick another copy of it (candidates of
		# one colour have mathematically equal potentials and BLAS rounds their rows differently), so the
		# comparison is on the seed VALUES, which is all the Lloyd run sees
		assert np.array_equal(X[a], X[b])
		assert np.array_equal(c, X[b])
		if name != "fewcolors":
			assert np.array_equal(a, b)
	hsva = eng.rgba_to_hsv(d)
	hpx, _ = eng.select_compact(hsva, 1, 30)
	lut = csp.hsv_feature_luts64()
	hr = hpx.cpu().numpy()
	F = np.stack([lut[c][hr[:, c]] for c in range(3)], axis=1)
	idx, cents = eng.kmeanspp_seeds(hpx, lut, min(k, 6), 10)
	ref = cs._seed_kmeans_plusplus(F, min(k, 6), 10)
	for a, b in zip(idx, ref):
		assert np.array_equal(F[a], F[b])


def test_host_register_in_place_and_pinned_outputs():
	"""a caller-owned image buffer page-locked in place gives the same results (the buffer is
	neither copied nor modified), and the output image comes back in page-locked memory as an ordinary,
	writable, C-contiguous HxWx4 uint8 array."""
	from image_segmenter_b200 import color_simplify as cs
	from image_segmenter_b200.engine import get_engine

	rgba = np.ascontiguousarray(blobby_rgba(5, 300, 400))
	before = rgba.copy()
	ref_out, ref_pal = cs.simplify_colors_median_cut(rgba, 16)
	eng = get_engine(0)
	eng.pin(rgba)
	try:
		out, pal = cs.simplify_colors_median_cut(rgba, 16)
		out2, pal2 = cs.simplify_colors_threshold(rgba, 16)
	finally:
		eng.unpin(rgba)
	assert np.array_equal(rgba, before)
	assert np.array_equal(out, ref_out) and np.array_equal(pal, ref_pal)
	assert out.flags["C_CONTIGUOUS"] and out.flags["WRITEABLE"] and out.dtype == np.uint8 and out.shape == rgba.shape
	out[0, 0, 0] ^= 1  # ordinary memory from the caller's point of view
	ro, rp = op.threshold(rgba, 16)
	assert np.array_equal(out2, ro) and np.array_equal(pal2, rp)
	with pytest.raises(ValueError):
		eng.pin(rgba[:, ::2])


# ---- degenerate and ragged inputs against the unmodified reference (tests/golden/reference_edge_cases.npz) ----
EDGE_IMAGES = ("tiny", "onepx", "ragged", "transparent", "dark", "midbright", "twocolors")


@pytest.fixture(scope="module")
def edge():
	from pathlib import Path

	return np.load(Path(__file__).resolve().parent / "golden" / "reference_edge_cases.npz", allow_pickle=False)


@pytest.mark.parametrize("name", EDGE_IMAGES)
def test_edge_cases_integer_and_palette_paths_bit_exact(edge, cs, name):
	"""1x1, 3x5, mixed alpha, fully transparent, all dark, fallback-threshold and two-colour images through
	the deterministic entry points: image, palette, dtype and the "same object back" early-outs as the reference."""
	img, cp = edge[f"in_{name}"], edge["custom_palette_in"]
	before = img.copy()
	calls = {
		"median_cut_8": lambda: cs.simplify_colors_median_cut(img, 8),
		"octree_5": lambda: cs.simplify_colors_octree(img, 5),
		"threshold_8": lambda: cs.simplify_colors_threshold(img, 8),
		"threshold_8_noalpha": lambda: cs.simplify_colors_threshold(img, 8, preserve_alpha=False),
		"custom_rgb": lambda: cs.simplify_colors_custom_palette(img, cp, True, "rgb"),
		"custom_lab": lambda: cs.simplify_colors_custom_palette(img, cp, True, "lab"),
		"custom_hsv_noalpha": lambda: cs.simplify_colors_custom_palette(img, cp, False, "hsv"),
	}
	for tag, fn in calls.items():
		out, pal = fn()
		key = f"{name}__{tag}"
		_eq(edge, key, out, pal)
		assert (out is img) == bool(edge[f"{key}__same_object"][0]), key
	assert np.array_equal(img, before)
	st = cs.get_color_statistics(img)
	ref = edge[f"{name}__stats"]
	assert st["total_unique_colors"] == int(ref[0]) and st["non_transparent_pixels"] == int(ref[1])
	assert np.allclose(st["rgb_mean"], ref[2:5], rtol=1e-12, atol=1e-12)
	assert np.allclose(st["rgb_std"], ref[5:8], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", EDGE_IMAGES)
def test_edge_cases_clustering_paths_shape_and_early_outs(edge, cs, name):
	"""The clustering entry points on the same inputs: the early-outs hand back the INPUT object with the
	reference's palette; otherwise shape, dtypes, alpha channel and the number of palette rows are the
	reference's (the colours themselves are subject to the documented tie rule on such tiny images)."""
	img = edge[f"in_{name}"]
	calls = {
		"kmeans_8": lambda: cs.simplify_colors_kmeans(img, 8, strict_reference_quirks=True),
		"hsv_4": lambda: cs.simplify_colors_hsv_clustering(img, 4),
		"perceptual_fast_4": lambda: cs.simplify_colors_perceptual_fast(img, 4),
		"perceptual_3": lambda: cs.simplify_colors_perceptual(img, 3, max_samples=2000),
	}
	for tag, fn in calls.items():
		key = f"{name}__{tag}"
		with warnings.catch_warnings():
			warnings.simplefilter("ignore")
			np.random.seed(7)
			out, pal = fn()
		ref_out, ref_pal = edge[f"{key}__rgba"], edge[f"{key}__palette"]
		if bool(edge[f"{key}__same_object"][0]):
			assert out is img, key
			assert np.array_equal(np.asarray(pal), ref_pal) and np.asarray(pal).dtype == ref_pal.dtype, key
			continue
		assert out is not img and out.shape == ref_out.shape and out.dtype == np.uint8, key
		assert np.array_equal(out[:, :, 3], ref_out[:, :, 3]), key
		assert np.asarray(pal).shape == ref_pal.shape and np.asarray(pal).dtype == ref_pal.dtype, key
		# ... and the colours.  The perceptual paths fit their palette on the host exactly as
		# the reference does and map the pixels with K4: bit-equal.  The k-means paths run Lloyd on the device:
		# equal to the reference unless an exact distance tie was decided the other way (DESIGN.md deviation 1),
		# in which case the palette must be the documented-rule replay's / of the same quality.
		if tag.startswith("perceptual"):
			assert np.array_equal(out, ref_out) and np.array_equal(np.asarray(pal), ref_pal), key
		elif tag == "kmeans_8":
			assert np.array_equal(out, ref_out), key  # strict quirk: RGB all zero + the alpha epilogue
			_kmeans_palette_ok(img, pal, ref_pal, key)
		else:
			_hsv_result_ok(img, out, pal, ref_out, ref_pal, key)


def test_host_buffer_call_equals_device_path():
	"""cs_host_lab_kmeans (host buffers in and out: chunked upload overlapped with the LAB conversion, batches of
	chained iterations, final E-step, download) returns exactly what the device-pointer path returns from the
	same image and initial centres: centres, labels, iteration count and inertia."""
	import ctypes as C

	from image_segmenter_b200 import _colorspace, _ffi
	from image_segmenter_b200.engine import KMeansGPU, get_engine
	from oracle import lab as olab

	eng = get_engine(0)
	rng = np.random.default_rng(77)
	h, w, K = 123, 77, 6  # 9471 px: not a multiple of 4 or of the 8 upload chunks
	rgba = np.dstack([rng.integers(0, 256, (h, w, 3), dtype=np.uint8), np.full((h, w), 255, np.uint8)])
	flat = np.ascontiguousarray(rgba.reshape(-1, 4))
	n = flat.shape[0]
	C0 = np.ascontiguousarray(olab.rgb2lab(flat[rng.choice(n, K, replace=False), :3]), dtype=np.float64)
	lut = np.ascontiguousarray(_colorspace.linear_lut256(), dtype=np.float64)
	cen = C0.copy()
	labels = np.full(n, 99, np.uint8)
	nit, inert = C.c_int(0), C.c_double(0.0)
	_ffi.check(eng.ctx.lib.cs_host_lab_kmeans(eng.ctx.handle, flat.ctypes.data, n, lut.ctypes.data, cen.ctypes.data, K, 15, 0.0,
	                                          _ffi.CS_LLOYD_EXACT_TIES, labels.ctypes.data, C.byref(nit), C.byref(inert)),
	           "cs_host_lab_kmeans")
	planes = eng.rgba_to_lab(eng.upload_rgba(rgba))
	fit = KMeansGPU(eng, "f32", n, planes=planes, exact=True).fit_single(C0, max_iter=15, tol=0.0)
	assert nit.value == fit.n_iter
	assert np.array_equal(cen, fit.centers)
	assert np.array_equal(labels, fit.labels[:n].cpu().numpy())
	assert abs(inert.value - fit.inertia) <= 1e-9 * max(1.0, fit.inertia)


# ---- BASELINE config 1 on its real input (VERDICT r1 missing #4) --------------------------------------------
@pytest.fixture(scope="module")
def working_image():
	"""app/working_image_cleaned.bmp of the reference (1024 x 1024, 9 colours, opaque), stored as a colour table
	and an index map (tests/golden/working_image_cleaned.npz, written by oracle/make_golden.py)."""
	from pathlib import Path

	g = np.load(Path(__file__).parent / "golden" / "working_image_cleaned.npz")
	rgb = g["colours"][g["index"]]
	return np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)])


def test_config1_working_image_vs_reference_known_answers(golden, cs, working_image):
	"""The five calls whose answers the unmodified reference gave on this image (bmp__* in
	reference_entry_points.npz): K collapses from 16 to the 7 colours that pass the brightness filter."""
	img = working_image
	true_cols = np.unique(img[:, :, :3].reshape(-1, 3), axis=0)
	assert len(true_cols) == 9
	out, pal = cs.simplify_colors_kmeans(img, 16)
	ref = golden["bmp__kmeans_16__palette"]
	assert pal.shape == ref.shape == (7, 3) and pal.dtype == np.uint8
	# every cluster is ONE colour: the exact mean is that colour.  The reference's float mean of identical values
	# can come out as 153.99999999999997 and truncate one too low (DESIGN.md deviation 2): equal or +1.
	d = pal.astype(int) - ref.astype(int)
	assert ((d == 0) | (d == 1)).all(), (pal, ref)
	assert {tuple(r) for r in pal} <= {tuple(r) for r in true_cols}
	bright = img[:, :, :3].astype(int).sum(2) > 90
	assert (out[~bright][:, :3] == 0).all() and np.array_equal(out[bright][:, :3], img[bright][:, :3])
	assert np.array_equal(out[:, :, 3], img[:, :, 3])
	out, pal = cs.simplify_colors_median_cut(img, 16)
	assert np.array_equal(pal, golden["bmp__median_cut_16__palette"]) and pal.dtype == np.int64
	assert np.array_equal(out, img)  # 9 colours, 16 boxes: every colour keeps its own box
	out, pal = cs.simplify_colors_threshold(img, 16)
	assert np.array_equal(pal, golden["bmp__threshold_16__palette"])
	out, pal = cs.simplify_colors_hsv_clustering(img, 16)
	assert np.array_equal(pal, golden["bmp__hsv_16__palette"])
	st = cs.get_color_statistics(img)
	ref = golden["bmp__stats"]
	assert st["total_unique_colors"] == int(ref[0]) and int(st["non_transparent_pixels"]) == int(ref[1])
	assert np.allclose(st["rgb_mean"], ref[2:5], rtol=1e-10) and np.allclose(st["rgb_std"], ref[5:8], rtol=1e-10)
