"""Drop-in for `app/processing/color_simplify.py` of jeffreyperez1620/image_segmenter.

Same 14 module-level names, positional order, defaults, return types and error messages as the
reference; every per-pixel step runs in the sm_100a kernels of
libcolorsimplify.so (include/colorsimplify.h) through `engine.Engine`.  What stays on the host
is what the reference also does on a palette-sized set: the global-RNG sampling of <= 10 000
colours, `np.unique` of those samples, Ward / KMeans fits of <= 5 000 sampled colours
(scikit-learn, exactly the reference's calls), the RandomState(42) stream and per-round decisions of
k-means++ seeding (every O(N) step of it is a kernel, engine.Engine.kmeanspp_seeds), and the
median-cut box tree over <= 65 536 histogram cells (C++, csrc/mediancut.cpp).

There is NO CPU fallback: without the CUDA library or a B200 every compute call raises
`_ffi.ColorSimplifyError` (a RuntimeError).

New keyword-only options (defaults keep the reference's behaviour unless stated):
  strict_reference_quirks   simplify_colors_kmeans: True reproduces the reference's no-op remap
                            (color_simplify.py:90 assigns into a temporary, so its RGB output is
                            all zero); the default False writes centres[labels] as intended.
  init_centers / max_iter / n_init / tol   injected initialisation and loop control for tests and
                            benchmarks (array init => one run, as sklearn does).
  fit="full" / process_group   simplify_colors_perceptual_fast: fit the LAB palette on every pixel on
                            the device (the headline kernel), optionally row-sharded over one process per GPU.
"""
from __future__ import annotations

import warnings
from fractions import Fraction
from typing import List, Tuple

import numpy as np

from . import _colorspace as cspace
from . import _ffi
from .engine import KMeansGPU, SampleKMeans, get_engine

__all__ = [
	"simplify_colors_kmeans", "simplify_colors_median_cut", "simplify_colors_octree", "simplify_colors_threshold",
	"simplify_colors_adaptive", "get_color_statistics", "simplify_colors_perceptual", "simplify_colors_perceptual_fast",
	"simplify_colors_adaptive_distance", "simplify_colors_hsv_clustering", "simplify_colors_custom_palette",
	"create_palette_from_colors", "check_gpu_availability", "get_recommended_algorithm",
]

# module-wide default for `strict_reference_quirks` (a caller that cannot pass keywords, like the
# reference's UI, can flip it here)
STRICT_REFERENCE_QUIRKS = False

_HSV_X2MAX = 2.0 ** 2 + 1.5 ** 2 + 1.0 ** 2  # bound on |feature|^2 of the weighted HSV rows


def _check_rgba(rgba) -> None:
	# color_simplify.py:34-35 (identical in every entry point)
	if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4:
		raise ValueError("rgba must be HxWx4 uint8")


def _check_k(K: int, what: str = "num_colors") -> None:
	"""The device clustering kernels hold at most CS_MAX_K = 256 centres (labels are one byte).  The reference
	accepts more; say so up front instead of failing deep inside a kernel call."""
	if K > _ffi.CS_MAX_K:
		raise ValueError(f"{what} resolves to {K} clusters; this B200 implementation supports at most {_ffi.CS_MAX_K}")


def _degenerate(rgba):
	# color_simplify.py:45-47, 72-74, 432-434, ...: the INPUT object and an int64 [[0,0,0]]
	return rgba, np.array([[0, 0, 0]])


# one-entry upload cache: simplify_colors_adaptive("adaptive") looks at the image twice (statistics, then the
# chosen algorithm) — the second call reuses the device copy instead of crossing PCIe again
_upload_cache = None


def _upload(eng, rgba):
	if _upload_cache is not None and _upload_cache[0] is rgba:
		return _upload_cache[1]
	return eng.upload_rgba(rgba)


def _download(d_out, shape) -> np.ndarray:
	"""Device (n,4) uint8 -> a NEW HxWx4 array (np.dstack in the reference).  The array lives in page-locked
	memory from torch's caching host allocator, so the copy is one DMA at PCIe rate instead of the driver's
	staged pageable copy (after the kernels the copies dominate); the block returns to the
	cache when the caller drops the array."""
	return get_engine().to_host(d_out).reshape(shape[0], shape[1], 4)


def _brightness_threshold(n_hi: int, n_lo: int, num_colors: int, hi: int, lo: int):
	"""The shared dark-pixel filter (color_simplify.py:56-64, 956-963): `> hi`, relaxed to `> lo`
	when fewer than num_colors pixels pass, relaxed to everything when none passes.
	Returns the integer threshold for the kernels (-1 = keep all)."""
	thr, cnt = hi, n_hi
	if cnt < num_colors:
		thr, cnt = lo, n_lo
	if cnt == 0:
		thr = -1
	return thr


def _moments_from_hist(hist: np.ndarray, lut3: np.ndarray):
	"""Exact mean / population variance of the three feature columns from per-byte counts
	(every feature is a function of one byte).  Rational arithmetic, rounded once — within
	1 ulp-ish of np.mean / np.var over the rows (sklearn/cluster/_kmeans.py:285-293)."""
	mean, var = np.zeros(3), np.zeros(3)
	for c in range(3):
		nz = np.nonzero(hist[c])[0]
		n = int(hist[c].sum())
		if n == 0:
			continue
		vals = [Fraction(float(lut3[c, v])) for v in nz]
		cnts = [int(hist[c, v]) for v in nz]
		s1 = sum(v * k for v, k in zip(vals, cnts))
		s2 = sum(v * v * k for v, k in zip(vals, cnts))
		m = s1 / n
		mean[c] = float(m)
		var[c] = float(s2 / n - m * m)
	return mean, var


def _seed_kmeans_plusplus(X: np.ndarray, K: int, n_init: int, seed: int = 42) -> List[np.ndarray]:
	"""The `n_init` k-means++ initialisations KMeans(random_state=seed).fit draws, as row indices
	into X.  Lloyd consumes no randomness, so drawing them back to back from one RandomState
	reproduces sklearn's stream (sklearn/cluster/_kmeans.py:1463-1514, 180-278).  HOST comparator for
	the tests only: the product seeds on the device (engine.Engine.kmeanspp_seeds)."""
	from sklearn.cluster._kmeans import _kmeans_plusplus
	from sklearn.utils import check_random_state
	from sklearn.utils.extmath import row_norms

	X = np.ascontiguousarray(X, dtype=np.float64)
	Xc = X - X.mean(axis=0)
	norms = row_norms(Xc, squared=True)
	w = np.ones(X.shape[0], dtype=np.float64)
	rs = check_random_state(seed)
	return [_kmeans_plusplus(Xc, K, norms, w, rs)[1] for _ in range(n_init)]


def _truncate_u8(centers: np.ndarray) -> np.ndarray:
	# np.clip(...).astype(np.uint8): truncation, never rounding (color_simplify.py:84, 534, 1002)
	return np.clip(centers, 0, 255).astype(np.uint8)


# =====================================================================================
def simplify_colors_kmeans(rgba: np.ndarray, num_colors: int = 8, preserve_alpha: bool = True, *,
                           strict_reference_quirks: bool | None = None, init_centers=None, max_iter: int = 300,
                           n_init: int = 10, tol: float | None = None) -> Tuple[np.ndarray, np.ndarray]:
	"""RGB k-means on the opaque, non-black pixels (reference color_simplify.py:12-102).

	Device steps: mask counts + distinct-colour bitmap (K cap, :44-70), Lloyd iterations on the
	packed RGBA8 pixels with exact integer sums (KMeans.fit, :79-80), label gather (:90) and alpha
	epilogue (:93-100).  Returns (HxWx4 uint8, K x 3 uint8 truncated centres)."""
	_check_rgba(rgba)
	eng = get_engine()
	d = _upload(eng, rgba)
	n_op, n_hi, n_lo, _ = eng.mask_stats(d, -1)
	if n_op == 0:
		return _degenerate(rgba)
	thr = _brightness_threshold(n_hi, n_lo, num_colors, 90, 30)  # mean(rgb) > 30 | 10  <=>  r+g+b > 90 | 30
	n_unique = eng.mask_stats(d, thr, want_unique=True)[3]
	K = min(int(num_colors), n_unique)
	if K < 2:
		return _degenerate(rgba)
	_check_k(K)
	ident = np.tile(np.arange(256, dtype=np.float64), (3, 1))
	_, var = _moments_from_hist(eng.channel_hist(d, 0, thr), ident)
	if tol is None:
		tol = float(np.mean(var) * 1e-4)  # _tolerance, sklearn/cluster/_kmeans.py:285-293
	if init_centers is not None:
		inits = [np.asarray(init_centers, dtype=np.float64).reshape(K, 3)]
	else:
		px, _ = eng.select_compact(d, 0, thr)  # rows of the reference's rgb_filtered, in its order
		_, inits = eng.kmeanspp_seeds(px, ident, K, n_init)
	km = KMeansGPU(eng, "rgba8", d.shape[0], px=d, mask_mode=0, min_bright=thr)
	fit = km.fit_best(inits, max_iter=max_iter, tol=tol)
	centers = _truncate_u8(fit.centers)
	# The last M-step's sums are exact integers, so the truncated mean can be taken exactly: floor(sum / count).
	# (A float mean that is an exact integer can land at 153.99999999999997 and truncate one too low — SURVEY.md
	# §0.3; the exact floor never does, and agrees with the float truncation everywhere else.)
	nz = fit.counts > 0
	centers[nz] = (fit.sums[nz].astype(np.int64) // fit.counts[nz].astype(np.int64)[:, None]).astype(np.uint8)
	quirk = STRICT_REFERENCE_QUIRKS if strict_reference_quirks is None else strict_reference_quirks
	pal = np.zeros_like(centers) if quirk else centers
	out = eng.remap_labels(d, fit.labels, pal, preserve_alpha, sel=(d, 0, thr))
	return _download(out, rgba.shape), centers


def _median_cut(rgba, num_colors, preserve_alpha):
	eng = get_engine()
	d = _upload(eng, rgba)
	out, pal, _ = eng.median_cut(d, int(num_colors), preserve_alpha)
	# getpalette()[:K] as a Python-int array: int64, possibly fewer than K rows (:148-149)
	return _download(out, rgba.shape), pal.astype(np.int64)[:num_colors]


def simplify_colors_median_cut(rgba: np.ndarray, num_colors: int = 8,
                               preserve_alpha: bool = True) -> Tuple[np.ndarray, np.ndarray]:
	"""Pillow MEDIANCUT on the RGB of every pixel, alpha ignored (reference :105-164): device
	histogram -> host box tree -> device box means -> device nearest-palette map, bit-exact."""
	_check_rgba(rgba)
	num_colors = 2 ** int(np.log2(num_colors))  # :131
	return _median_cut(rgba, num_colors, preserve_alpha)


def simplify_colors_octree(rgba: np.ndarray, num_colors: int = 8,
                           preserve_alpha: bool = True) -> Tuple[np.ndarray, np.ndarray]:
	"""The reference's "octree" also asks Pillow for MEDIANCUT (:201) — median cut without the
	power-of-two rounding (reference :167-220)."""
	_check_rgba(rgba)
	return _median_cut(rgba, num_colors, preserve_alpha)


def simplify_colors_threshold(rgba: np.ndarray, num_colors: int = 8,
                              preserve_alpha: bool = True) -> Tuple[np.ndarray, np.ndarray]:
	"""Posterize: (c // step) * step per channel; palette = first K rows of the sorted distinct
	output colours (reference :223-277).  One device pass writes the image and marks the colours."""
	_check_rgba(rgba)
	levels = int(np.ceil(np.cbrt(num_colors)))  # host, as the reference (:255-256)
	step = 256 // levels
	eng = get_engine()
	d = _upload(eng, rgba)
	out, uniq = eng.posterize(d, step, preserve_alpha)
	return _download(out, rgba.shape), uniq[:num_colors]


def simplify_colors_adaptive(rgba: np.ndarray, target_colors: int = 8, preserve_alpha: bool = True,
                             algorithm: str = "kmeans") -> Tuple[np.ndarray, np.ndarray]:
	"""String dispatch of the reference (:280-342), including the "adaptive" heuristic and the
	default-to-kmeans branch."""
	table = {
		"kmeans": simplify_colors_kmeans, "median_cut": simplify_colors_median_cut, "octree": simplify_colors_octree,
		"threshold": simplify_colors_threshold, "perceptual": simplify_colors_perceptual,
		"perceptual_fast": simplify_colors_perceptual_fast, "adaptive_distance": simplify_colors_adaptive_distance,
		"hsv_clustering": simplify_colors_hsv_clustering,
	}
	if algorithm in table:
		return table[algorithm](rgba, target_colors, preserve_alpha)
	if algorithm == "custom_palette":
		raise ValueError("Custom palette requires palette parameter")
	if algorithm == "adaptive":
		global _upload_cache
		_check_rgba(rgba)
		_upload_cache = (rgba, get_engine().upload_rgba(rgba))
		try:
			total = get_color_statistics(rgba)["total_unique_colors"]
			if total <= target_colors:
				return simplify_colors_threshold(rgba, target_colors, preserve_alpha)
			if total > 1000:
				return simplify_colors_perceptual(rgba, target_colors, preserve_alpha)
			return simplify_colors_hsv_clustering(rgba, target_colors, preserve_alpha)
		finally:
			_upload_cache = None
	return simplify_colors_kmeans(rgba, target_colors, preserve_alpha)


def get_color_statistics(rgba: np.ndarray) -> dict:
	"""Distinct RGBA count, opaque count, fp64 mean / population std of RGB over alpha > 0
	(reference :345-384).  One device pass: 2^32-bit presence bitmap + exact u64 moments; the
	mean / std are formed from the exact integer sums (rational, rounded once)."""
	_check_rgba(rgba)
	eng = get_engine()
	d = _upload(eng, rgba)
	n_unique, n_op, s1, s2 = eng.statistics(d)
	if n_op > 0:
		mean = np.array([float(Fraction(s, n_op)) for s in s1])
		var = [Fraction(q, n_op) - Fraction(s, n_op) ** 2 for s, q in zip(s1, s2)]
		std = np.sqrt(np.array([float(v) for v in var]))
	else:
		mean, std = np.array([0, 0, 0]), np.array([0, 0, 0])
	return {
		"total_unique_colors": int(n_unique),
		"non_transparent_pixels": np.int64(n_op),
		"rgb_mean": mean,
		"rgb_std": std,
		"image_size": rgba.shape[:2],
	}


def _sample_opaque(eng, d, n_op: int, max_samples: int) -> np.ndarray:
	"""rgb_flat (opaque pixels, scan order), or rgb_flat[np.random.choice(len, m, replace=False)]
	with the GLOBAL NumPy RNG exactly as the reference draws it (:442-448)."""
	src = d if n_op == d.shape[0] else eng.select_compact(d, 0, -1)[0]
	if n_op > max_samples:
		indices = np.random.choice(n_op, max_samples, replace=False)
		return eng.gather(src, indices).cpu().numpy()[:, :3]
	return src.cpu().numpy()[:, :3]


def _filter_dark_unique(unique_colors: np.ndarray, num_colors: int) -> np.ndarray:
	"""Boolean keep-mask over the (palette-sized) unique colours (:455-463, 644-652)."""
	brightness = np.mean(unique_colors, axis=1)
	keep = brightness > 30
	if np.sum(keep) < num_colors:
		keep = brightness > 10
	if np.sum(keep) == 0:
		keep = np.ones(len(unique_colors), dtype=bool)
	return keep


def simplify_colors_perceptual(rgba: np.ndarray, num_colors: int = 8, preserve_alpha: bool = True,
                               color_tolerance: float = 30.0, use_gpu: bool = False,
                               max_samples: int = 10000) -> Tuple[np.ndarray, np.ndarray]:
	"""Ward clustering of <= max_samples sampled colours in LAB, then the full-image nearest-centre
	remap (reference :387-559).  The palette fit is host/sklearn as in the reference; the full-image
	LAB conversion + argmin + gather (:540-547) is one fused device kernel.  Keeps the reference's
	quirk: the search compares LAB pixels with the RGB-valued uint8 centres.  `color_tolerance` and
	`use_gpu` are accepted and unused (the reference's use_gpu branch computes the same thing)."""
	_check_rgba(rgba)
	eng = get_engine()
	d = _upload(eng, rgba)
	n_op = eng.mask_stats(d, -1)[0]
	if n_op == 0:
		return _degenerate(rgba)
	samples = _sample_opaque(eng, d, n_op, max_samples)
	unique_colors, counts = np.unique(samples, axis=0, return_counts=True)
	keep = _filter_dark_unique(unique_colors, num_colors)
	uniq, cnt = unique_colors[keep], counts[keep]
	K = min(num_colors, len(uniq))
	if K < 2:
		return _degenerate(rgba)
	_check_k(K)
	from sklearn.cluster import AgglomerativeClustering

	lab = cspace.rgb2lab_small(uniq)
	clustering = AgglomerativeClustering(n_clusters=K, linkage="ward", distance_threshold=None)
	labels = clustering.fit_predict(lab)
	centers = np.zeros((clustering.n_clusters_, 3))
	for i in range(clustering.n_clusters_):
		m = labels == i
		if np.any(m):
			centers[i] = np.average(uniq[m], weights=cnt[m], axis=0)
	centers = _truncate_u8(centers)
	out, _ = eng.assign_remap(d, _ffi.CS_SPACE_LAB, centers.astype(np.float64), centers, preserve_alpha)
	return _download(out, rgba.shape), centers


def simplify_colors_perceptual_fast(rgba: np.ndarray, num_colors: int = 8, preserve_alpha: bool = True,
                                    color_tolerance: float = 30.0, *, fit: str = "sample", init_centers=None,
                                    max_iter: int = 100, tol: float | None = None,
                                    process_group=None) -> Tuple[np.ndarray, np.ndarray]:
	"""LAB k-means palette from a <= 512 px, <= 5000-colour sample, then the full-image LAB
	nearest-centre remap (reference :562-707).  Downsample / sample / fit are host-sized and follow
	the reference call for call (cv.resize INTER_AREA, global-RNG choice, sklearn KMeans); the
	per-pixel tail (:688-695) is the fused device kernel.

	Keyword-only additions (defaults = the reference's behaviour):
	  fit="full"       the palette is fitted on EVERY opaque pixel on the device: fp32 CIELAB planes (K1),
	                   then Lloyd iterations (K2/K3, exact labels) from `init_centers` (K x 3 CIELAB) or, when
	                   None, from the centres of the reference's own sample fit — `max_iter` iterations at
	                   most, stop at sum(shift^2) <= `tol` (None: sklearn's 1e-4 * mean feature variance of
	                   the sample).  This is the path bench.py measures (BASELINE metric).
	  process_group    a torch.distributed group (one process per GPU): `rgba` is THIS rank's block of rows of
	                   a row-sharded image; the ranks fit one common palette (centre partials exchanged every
	                   iteration, image_segmenter_b200.sharded) and each returns its own rows.  Needs fit="full"
	                   and `init_centers` (the sample fit would differ between ranks)."""
	import cv2 as cv

	_check_rgba(rgba)
	if fit not in ("sample", "full"):
		raise ValueError("fit must be 'sample' or 'full'")
	if process_group is not None and (fit != "full" or init_centers is None):
		raise ValueError("process_group needs fit='full' and init_centers")
	h, w = rgba.shape[:2]
	eng = get_engine()
	planes = None
	if fit == "full":
		d, planes = eng.upload_rgba_lab(rgba)
	else:
		d = _upload(eng, rgba)
	n_op = eng.mask_stats(d, -1)[0]
	if n_op == 0 and process_group is None:
		return _degenerate(rgba)
	centers_lab = None
	if init_centers is not None:
		centers_lab = np.ascontiguousarray(init_centers, dtype=np.float64).reshape(-1, 3)
		K = centers_lab.shape[0]
		if not 1 <= K <= _ffi.CS_MAX_K:
			raise ValueError(f"init_centers must have between 1 and {_ffi.CS_MAX_K} rows")
	else:
		max_dim = 512
		if h > max_dim or w > max_dim:
			scale = min(max_dim / h, max_dim / w)
			new_h, new_w = int(h * scale), int(w * scale)
			# one 4-channel INTER_AREA pass instead of the reference's two (:608-614): the channels are resized
			# independently, so the planes equal cv.resize(rgb) / cv.resize(alpha) bit for bit (tests/test_host_logic.py),
			# without the 3-of-4-channel gather copy
			small = cv.resize(np.ascontiguousarray(rgba), (new_w, new_h), interpolation=cv.INTER_AREA)
			rgb_small, alpha_small = small[:, :, :3], small[:, :, 3]
			nts = alpha_small > 0
			if not np.any(nts):
				return _degenerate(rgba)
			rgb_flat = rgb_small[nts].reshape(-1, 3)
		else:
			rgb_flat = rgba[:, :, :3][rgba[:, :, 3] > 0].reshape(-1, 3)  # <= 512 x 512: sample-sized
		sample_size = min(5000, len(rgb_flat))
		if len(rgb_flat) > sample_size:
			rgb_flat = rgb_flat[np.random.choice(len(rgb_flat), sample_size, replace=False)]
		unique_colors = np.unique(rgb_flat, axis=0)
		uniq = unique_colors[_filter_dark_unique(unique_colors, num_colors)]
		K = min(num_colors, len(uniq))
		if K < 2:
			return _degenerate(rgba)
		_check_k(K)
		lab = cspace.rgb2lab_small(uniq)
		# KMeans(n_clusters=K, random_state=42, n_init=10, max_iter=100).fit(lab) (:669-675) on the device: the ten
		# k-means++ seedings in lockstep, then the ten Lloyd loops in one launch (engine.SampleKMeans)
		km = SampleKMeans(eng).fit(lab, K, n_init=10, max_iter=100, seed=42)
		centers_lab = km.cluster_centers_
		if tol is None:
			tol = float(np.mean(np.var(lab, axis=0)) * 1e-4)
	if fit == "full":
		centers_lab = _fit_lab_full(eng, d, planes, n_op, centers_lab, int(max_iter), 0.0 if tol is None else float(tol),
		                            process_group)
	centers_rgb = _truncate_u8(cspace.lab2rgb_small(centers_lab) * 255)
	out, _ = eng.assign_remap(d, _ffi.CS_SPACE_LAB, centers_lab, centers_rgb, preserve_alpha)
	return _download(out, rgba.shape), centers_rgb


def _fit_lab_full(eng, d, planes, n_op: int, centers: np.ndarray, max_iter: int, tol: float, group) -> np.ndarray:
	"""Lloyd iterations over every opaque pixel of the (local) image in fp32 CIELAB (KMeans.fit's loop,
	sklearn/cluster/_kmeans.py:705-738, on the rows the reference would have fitted had it not sampled)."""
	n = d.shape[0]
	if n_op != n:
		# transparent pixels take no part in the fit (:628): fit on the compacted opaque pixels
		src = eng.select_compact(d, 0, -1)[0]
		planes = eng.rgba_to_lab(src)
		n = src.shape[0]
	if group is not None:
		from .sharded import make_gpu_lloyd

		drv = make_gpu_lloyd(eng, planes, n, centers.shape[0], group=group, exact=True)
		return drv.run(centers, max_iter, tol).centers
	km = KMeansGPU(eng, "f32", n, planes=planes, exact=True)
	return km.fit_centers(centers, max_iter=max_iter, tol=tol)


def simplify_colors_adaptive_distance(rgba: np.ndarray, num_colors: int = 8, preserve_alpha: bool = True,
                                      similarity_threshold: float = 25.0) -> Tuple[np.ndarray, np.ndarray]:
	"""DBSCAN on standardised LAB of every opaque pixel (reference :710-882).

	DBSCAN's neighbourhood queries are not a streaming per-pixel kernel and stay in scikit-learn,
	exactly the reference's calls; the device does the LAB conversion of every pixel, the full-N
	KMeans fallback fit (k-means++ passes + fp64 Lloyd on the standardised rows, engine.KMeansRows64),
	the nearest-clustered-pixel search for the dark pixels, the per-label RGB sums and the final gather.  The reference's "too many clusters" branch indexes a positional
	centre array with raw cluster ids (:841-846 vs :858, :870) and raises IndexError or paints wrong
	colours; that behaviour is kept (the labels are handed to the gather as they are, out-of-range
	ids raise IndexError like the reference)."""
	_check_rgba(rgba)
	eng = get_engine()
	d = _upload(eng, rgba)
	n_op = eng.mask_stats(d, -1)[0]
	if n_op == 0:
		return _degenerate(rgba)
	src, src_index = (d, None) if n_op == d.shape[0] else eng.select_compact(d, 0, -1, want_index=True)
	d_lab64 = eng.rgba_to_lab_f64(src)
	lab_flat = d_lab64.cpu().numpy()
	keep = lab_flat[:, 0] > 10
	if np.sum(keep) < num_colors:
		keep = lab_flat[:, 0] > 5
	if np.sum(keep) == 0:
		keep = np.ones(len(lab_flat), dtype=bool)
	lab_f = lab_flat[keep]
	from sklearn.cluster import DBSCAN
	from sklearn.preprocessing import StandardScaler

	lab_n = StandardScaler().fit_transform(lab_f)
	eps = (similarity_threshold / 100.0) * 0.5
	cl = DBSCAN(eps=eps, min_samples=3).fit_predict(lab_n)
	if -1 in cl:
		from sklearn.neighbors import NearestNeighbors

		noise, good = cl == -1, cl != -1
		if np.any(good):
			nn = NearestNeighbors(n_neighbors=1).fit(lab_n[good])
			cl[noise] = cl[good][nn.kneighbors(lab_n[noise])[1].flatten()]
	n_clusters = len(np.unique(cl))
	if n_clusters < num_colors:
		# the full-N fallback fit (:809-814) on the device: k-means++ passes + fp64 Lloyd on the standardised rows
		_check_k(int(num_colors))
		import torch

		from .engine import KMeansRows64

		d_rows = torch.from_numpy(np.ascontiguousarray(lab_n, dtype=np.float64)).to(eng.dev)
		cl = KMeansRows64(eng, d_rows).fit_predict(int(num_colors))
		n_clusters = num_colors
	if n_clusters > num_colors:
		sizes = np.bincount(cl.astype(int))
		order = np.argsort(sizes)
		keep_ids, merge_ids = order[-num_colors:], order[:-num_colors]
		for small in merge_ids:
			c_small = np.mean(lab_f[cl == small], axis=0)
			c_large = np.array([np.mean(lab_f[cl == big], axis=0) for big in keep_ids])
			cl[cl == small] = keep_ids[np.argmin(np.linalg.norm(c_large - c_small, axis=1))]
	uniq_labels = np.unique(cl)
	if uniq_labels.min() < 0:
		# every point was DBSCAN noise and stayed -1: the reference paints with centres[-1] through negative
		# indexing (:870); map the id to the last centre as NumPy would
		cl = np.where(cl < 0, len(uniq_labels) - 1, cl)
		uniq_labels = np.unique(cl)
	if uniq_labels.max() >= 255:
		raise ValueError(f"cluster id {int(uniq_labels.max())} does not fit the one-byte label map of the device gather "
		                 "(255 is the 'no label' value); the reference accepts such ids")
	# labels of every opaque pixel: clustered ones keep their id, dark ones take the id of the
	# nearest clustered pixel in LAB (:861-867)
	all_labels = np.zeros(len(lab_flat), dtype=np.int64)
	all_labels[np.where(keep)[0]] = cl
	dark = np.where(~keep)[0]
	if len(dark):
		# nearest clustered pixel in LAB for every dark pixel (:861-867), on the device
		import torch

		d_dark = d_lab64[torch.from_numpy(dark).to(eng.dev)]
		d_kept = d_lab64[torch.from_numpy(np.where(keep)[0]).to(eng.dev)]
		near = eng.nn_argmin_rows64(d_dark, d_kept)
		all_labels[dark] = cl[near]
	torch = __import__("torch")
	lab_u8 = np.full(d.shape[0], 255, dtype=np.uint8)
	if src_index is None:
		lab_u8[:] = all_labels
		fit_u8 = np.where(keep, all_labels, 255).astype(np.uint8)
	else:
		pos = src_index.cpu().numpy()
		lab_u8[pos] = all_labels
		fit_u8 = np.full(d.shape[0], 255, dtype=np.uint8)
		fit_u8[pos[keep]] = cl
	d_fit = torch.from_numpy(fit_u8).to(eng.dev)
	acc = eng.sum_by_label(d, d_fit, int(uniq_labels.max()) + 1)
	centers = np.zeros((len(uniq_labels), 3))
	for i, lab_id in enumerate(uniq_labels):  # positional centres (:841-846)
		centers[i] = acc[lab_id, :3] / acc[lab_id, 3]
	centers = _truncate_u8(centers)
	if all_labels.max() >= len(centers):
		raise IndexError(f"index {int(all_labels.max())} is out of bounds for axis 0 with size {len(centers)}")
	out = eng.remap_labels(d, torch.from_numpy(lab_u8).to(eng.dev), centers, preserve_alpha)
	return _download(out, rgba.shape), centers


def simplify_colors_hsv_clustering(rgba: np.ndarray, num_colors: int = 8, preserve_alpha: bool = True,
                                   hue_tolerance: float = 15.0, saturation_tolerance: float = 0.2, *,
                                   init_centers=None, max_iter: int = 300, n_init: int = 10,
                                   tol: float | None = None) -> Tuple[np.ndarray, np.ndarray]:
	"""k-means on weighted OpenCV-HSV features of the opaque, V > 30 pixels; centres are the RGB
	means of the clusters; dark pixels go to the nearest centre in RGB (reference :885-1036).
	All per-pixel steps on the device: RGB->HSV, mask counts + distinct-triple bitmap, Lloyd
	iterations through per-byte feature tables, per-label RGB sums, RGB argmin for the dark pixels,
	label merge and gather.  `hue_tolerance` / `saturation_tolerance` are accepted and unused, as in
	the reference."""
	_check_rgba(rgba)
	eng = get_engine()
	torch = __import__("torch")
	d = _upload(eng, rgba)
	hsva = eng.rgba_to_hsv(d)
	n_op, n_hi, n_lo, _ = eng.mask_stats(hsva, -1, hsv=True)
	if n_op == 0:
		return _degenerate(rgba)
	thr = _brightness_threshold(n_hi, n_lo, num_colors, 30, 10)
	n_unique = eng.mask_stats(hsva, thr, hsv=True, want_unique=True)[3]
	K = min(int(num_colors), n_unique)
	if K < 2:
		return _degenerate(rgba)
	_check_k(K)
	lut3 = cspace.hsv_feature_luts()
	if tol is None:
		_, var = _moments_from_hist(eng.channel_hist(hsva, 1, thr), lut3.astype(np.float64))
		tol = float(np.mean(var) * 1e-4)
	if init_centers is not None:
		inits = [np.asarray(init_centers, dtype=np.float64).reshape(K, 3)]
	else:
		px, _ = eng.select_compact(hsva, 1, thr)
		_, inits = eng.kmeanspp_seeds(px, cspace.hsv_feature_luts64(), K, n_init)
	d_lut = torch.from_numpy(lut3).to(eng.dev)
	km = KMeansGPU(eng, "px8lut", d.shape[0], px=hsva, lut3=d_lut, mask_mode=1, min_bright=thr, x2max=_HSV_X2MAX)
	fit = km.fit_best(inits, max_iter=max_iter, tol=tol)
	sel = (hsva, 1, thr)
	acc = eng.sum_by_label(d, fit.labels, K, sel=sel)
	centers = np.zeros((K, 3))
	nz = acc[:, 3] > 0
	centers[nz] = acc[nz, :3] / acc[nz, 3:4]  # np.mean of uint8 rows: exact integer sum / count
	centers = _truncate_u8(centers)
	# dark (filtered-out) opaque pixels -> nearest centre in RGB (:1015-1021)
	_, near = eng.assign_remap(d, _ffi.CS_SPACE_RGB, centers.astype(np.float64), centers, preserve_alpha, want_labels=True)
	merged = eng.merge_labels(fit.labels, near, d.shape[0], sel=sel)
	out = eng.remap_labels(d, merged, centers, preserve_alpha, sel=(d, 0, -1))
	return _download(out, rgba.shape), centers


def simplify_colors_custom_palette(rgba: np.ndarray, custom_palette: np.ndarray, preserve_alpha: bool = True,
                                   distance_metric: str = "lab") -> Tuple[np.ndarray, np.ndarray]:
	"""Map every opaque pixel to the nearest colour of a user palette in LAB, OpenCV-HSV (no hue
	wrap) or RGB (reference :1039-1123).  One fused device kernel (conversion + argmin + gather +
	alpha); the palette's own features are computed on the host (palette-sized)."""
	_check_rgba(rgba)
	if custom_palette.dtype != np.uint8 or custom_palette.ndim != 2 or custom_palette.shape[1] != 3:
		raise ValueError("custom_palette must be Nx3 uint8")
	if custom_palette.shape[0] < 1 or custom_palette.shape[0] > _ffi.CS_MAX_K:
		raise ValueError(f"custom_palette must have between 1 and {_ffi.CS_MAX_K} colours")
	eng = get_engine()
	d = _upload(eng, rgba)
	if eng.mask_stats(d, -1)[0] == 0:
		return rgba, custom_palette
	if distance_metric == "lab":
		space, feats = _ffi.CS_SPACE_LAB, cspace.rgb2lab_small(custom_palette)
	elif distance_metric == "hsv":
		space, feats = _ffi.CS_SPACE_HSV, cspace.rgb2hsv_u8_small(custom_palette).astype(np.float64)
	else:
		space, feats = _ffi.CS_SPACE_RGB, custom_palette.astype(np.float64)
	out, _ = eng.assign_remap(d, space, feats, custom_palette, preserve_alpha)
	return _download(out, rgba.shape), custom_palette


def create_palette_from_colors(colors: List[Tuple[int, int, int]]) -> np.ndarray:
	"""reference :1126-1141."""
	return np.array(colors, dtype=np.uint8)


def check_gpu_availability() -> dict:
	"""Same keys as the reference (:1144-1187); reports the CUDA devices this library would use."""
	info = {"cupy_available": False, "pytorch_available": False, "cuda_available": False, "gpu_count": 0,
	        "gpu_names": []}
	try:
		import torch

		info["pytorch_available"] = True
		if torch.cuda.is_available():
			info["cuda_available"] = True
			info["gpu_count"] = torch.cuda.device_count()
			info["gpu_names"] = [torch.cuda.get_device_name(i) for i in range(info["gpu_count"])]
	except ImportError:
		pass
	return info


def get_recommended_algorithm(image_size: tuple, gpu_available: bool = False) -> str:
	"""reference :1190-1219 (thresholds 1 MP / 500 K / 100 K pixels)."""
	h, w = image_size
	total = h * w
	if total > 1000000:
		return "perceptual" if gpu_available else "perceptual_fast"
	if total > 500000:
		return "perceptual_fast"
	if total > 100000:
		return "hsv_clustering"
	return "kmeans"
