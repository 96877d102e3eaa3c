"""End-to-end CPU restatement of the reference's colour-simplification entry points.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function follows the body of the
same-named `simplify_colors_*` function of app/processing/color_simplify.py (line ranges in the
docstrings) but is assembled from this package's restatements (lab.py, kmeans.py, hsv.py,
mediancut.py) plus the third-party fits the reference itself delegates to scikit-learn
(KMeans with k-means++ seeding, Ward agglomerative clustering).  It exists so that the GPU
box — where /root/reference does not exist — can check whole entry points, and it is pinned by
tests/golden/*.npz, which were produced by the UNMODIFIED reference module in the authoring
container (oracle/make_golden.py).

The shared epilogue (color_simplify.py:93-100 and the identical blocks of every function):
alpha_out = alpha, or (alpha > 128) * 255 when preserve_alpha is False; np.dstack.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import hsv as ohsv
from . import kmeans as okm
from . import lab as olab
from . import mediancut as omc


def _check(rgba):
	if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4:
		raise ValueError("rgba must be HxWx4 uint8")


def _epilogue(rgb_q, alpha, preserve_alpha):
	a = alpha if preserve_alpha else (alpha > 128).astype(np.uint8) * 255
	return np.dstack([rgb_q, a])


def _dark_filter(brightness, num_colors, hi=30, lo=10):
	"""The brightness filter shared by kmeans / perceptual / perceptual_fast / hsv
	(color_simplify.py:56-64, 455-463, 644-652, 956-963)."""
	m = brightness > hi
	if m.sum() < num_colors:
		m = brightness > lo
	if m.sum() == 0:
		m = np.ones(len(brightness), dtype=bool)
	return m


def posterize_step(num_colors: int) -> int:
	"""color_simplify.py:255-256."""
	levels = int(np.ceil(np.cbrt(num_colors)))
	return 256 // levels


def threshold(rgba, num_colors=8, preserve_alpha=True):
	"""simplify_colors_threshold (color_simplify.py:223-277)."""
	_check(rgba)
	step = posterize_step(num_colors)
	rgb = rgba[:, :, :3]
	q = (rgb // step) * step
	out = _epilogue(q, rgba[:, :, 3], preserve_alpha)
	palette = np.unique(q.reshape(-1, 3), axis=0)[:num_colors]
	return out, palette


def median_cut(rgba, num_colors=8, preserve_alpha=True, power_of_two=True):
	"""simplify_colors_median_cut (color_simplify.py:105-164); power_of_two=False gives
	simplify_colors_octree (:167-220), which also calls MEDIANCUT."""
	_check(rgba)
	if power_of_two:
		num_colors = 2 ** int(np.log2(num_colors))
	rgb = np.ascontiguousarray(rgba[:, :, :3])
	pal, idx = omc.quantize(rgb, num_colors)
	palette = pal.astype(np.int64)[:num_colors]
	return _epilogue(pal[idx], rgba[:, :, 3], preserve_alpha), palette


def statistics(rgba):
	"""get_color_statistics (color_simplify.py:345-384)."""
	_check(rgba)
	words = np.ascontiguousarray(rgba).view(np.uint32).reshape(-1)
	nt = rgba[:, :, 3] > 0
	n = int(nt.sum())
	if n > 0:
		sel = rgba[nt][:, :3]
		mean, std = np.mean(sel, axis=0), np.std(sel, axis=0)
	else:
		mean, std = np.array([0, 0, 0]), np.array([0, 0, 0])
	return {"total_unique_colors": int(len(np.unique(words))), "non_transparent_pixels": n,
	        "rgb_mean": mean, "rgb_std": std, "image_size": rgba.shape[:2]}


def kmeans_rgb(rgba, num_colors=8, preserve_alpha=True, intended_remap=True):
	"""simplify_colors_kmeans (color_simplify.py:12-102).  intended_remap=True writes
	centres[labels] into the filtered pixels (what line 90 means to do); False reproduces the
	reference's actual output, whose RGB is all zero because line 90 assigns into a temporary."""
	from sklearn.cluster import KMeans

	_check(rgba)
	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	if not nt.any():
		return rgba, np.array([[0, 0, 0]])
	flat = rgb[nt].reshape(-1, 3)
	keep = _dark_filter(np.mean(flat, axis=1), num_colors)
	filt = flat[keep]
	k = min(num_colors, len(np.unique(filt, axis=0)))
	if k < 2:
		return rgba, np.array([[0, 0, 0]])
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		km = KMeans(n_clusters=k, random_state=42, n_init=10)
		labels = km.fit_predict(filt)
	centers = np.clip(km.cluster_centers_, 0, 255).astype(np.uint8)
	q = np.zeros_like(rgb)
	if intended_remap:
		sub = np.zeros_like(flat)
		sub[keep] = centers[labels]
		q[nt] = sub
	return _epilogue(q, alpha, preserve_alpha), centers


def assign_remap(rgba, centers_feat, palette_rgb, space, preserve_alpha=True):
	"""The full-image tail shared by perceptual / perceptual_fast / custom_palette
	(color_simplify.py:537-557, 685-705, 1086-1121): features of every alpha>0 pixel, nearest
	centre (pairwise_distances_argmin_min), gather, epilogue.  Returns (rgba_out, indices)."""
	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	px = rgb[nt].reshape(-1, 3)
	if space == "lab":
		feat = olab.rgb2lab(px.reshape(-1, 1, 3)).reshape(-1, 3)
	elif space == "hsv":
		feat = ohsv.rgb_to_hsv_u8(px)
	else:
		feat = px
	idx, _ = okm.argmin_min(feat, centers_feat)
	q = np.zeros_like(rgb)
	q[nt] = np.asarray(palette_rgb)[idx]
	full_idx = np.full(alpha.shape, 255, dtype=np.uint8)
	full_idx[nt] = idx
	return _epilogue(q, alpha, preserve_alpha), full_idx


def custom_palette(rgba, palette, preserve_alpha=True, distance_metric="lab"):
	"""simplify_colors_custom_palette (color_simplify.py:1039-1123)."""
	_check(rgba)
	if palette.dtype != np.uint8 or palette.ndim != 2 or palette.shape[1] != 3:
		raise ValueError("custom_palette must be Nx3 uint8")
	if not (rgba[:, :, 3] > 0).any():
		return rgba, palette
	if distance_metric == "lab":
		cf = olab.rgb2lab(palette.reshape(-1, 1, 3)).reshape(-1, 3)
		sp = "lab"
	elif distance_metric == "hsv":
		cf = ohsv.rgb_to_hsv_u8(palette)
		sp = "hsv"
	else:
		cf, sp = palette, "rgb"
	out, _ = assign_remap(rgba, cf, palette, sp, preserve_alpha)
	return out, palette


def perceptual_fast_fit(rgba, num_colors=8):
	"""The palette fit of simplify_colors_perceptual_fast (color_simplify.py:593-682): INTER_AREA
	downsample to <= 512 px, <= 5000 samples drawn with the GLOBAL NumPy RNG, unique, brightness
	filter, LAB, KMeans(random_state=42, n_init=10, max_iter=100).  Returns (lab_centres fp64,
	rgb_centres uint8) or None when the reference returns early."""
	import cv2 as cv
	from sklearn.cluster import KMeans

	h, w = rgba.shape[:2]
	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	if not nt.any():
		return None
	if h > 512 or w > 512:
		scale = min(512 / h, 512 / w)
		nh, nw = int(h * scale), int(w * scale)
		rgb_s = cv.resize(rgb, (nw, nh), interpolation=cv.INTER_AREA)
		a_s = cv.resize(alpha, (nw, nh), interpolation=cv.INTER_AREA)
		nts = a_s > 0
		if not nts.any():
			return None
		flat = rgb_s[nts].reshape(-1, 3)
	else:
		flat = rgb[nt].reshape(-1, 3)
	n_s = min(5000, len(flat))
	if len(flat) > n_s:
		flat = flat[np.random.choice(len(flat), n_s, replace=False)]
	uniq = np.unique(flat, axis=0)
	uniq = uniq[_dark_filter(np.mean(uniq, axis=1), num_colors)]
	lab_u = olab.rgb2lab(uniq.reshape(-1, 1, 3)).reshape(-1, 3)
	k = min(num_colors, len(uniq))
	if k < 2:
		return None
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		km = KMeans(n_clusters=k, random_state=42, n_init=10, max_iter=100).fit(lab_u)
	cl = km.cluster_centers_
	crgb = np.clip(olab.lab2rgb(cl.reshape(-1, 1, 3)).reshape(-1, 3) * 255, 0, 255).astype(np.uint8)
	return cl, crgb


def perceptual_fast(rgba, num_colors=8, preserve_alpha=True):
	"""simplify_colors_perceptual_fast (color_simplify.py:562-707)."""
	_check(rgba)
	fit = perceptual_fast_fit(rgba, num_colors)
	if fit is None:
		return rgba, np.array([[0, 0, 0]])
	cl, crgb = fit
	out, _ = assign_remap(rgba, cl, crgb, "lab", preserve_alpha)
	return out, crgb


def perceptual_fit(rgba, num_colors=8, max_samples=10000):
	"""The palette fit of simplify_colors_perceptual (color_simplify.py:424-534): <= max_samples
	pixels drawn with the GLOBAL NumPy RNG, unique + counts, brightness filter, LAB, Ward
	agglomerative clustering, count-weighted RGB means truncated to uint8."""
	from sklearn.cluster import AgglomerativeClustering

	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	if not nt.any():
		return None
	flat = rgb[nt].reshape(-1, 3)
	if len(flat) > max_samples:
		flat = flat[np.random.choice(len(flat), max_samples, replace=False)]
	uniq, counts = np.unique(flat, axis=0, return_counts=True)
	keep = _dark_filter(np.mean(uniq, axis=1), num_colors)
	uniq, counts = uniq[keep], counts[keep]
	lab_u = olab.rgb2lab(uniq.reshape(-1, 1, 3)).reshape(-1, 3)
	k = min(num_colors, len(uniq))
	if k < 2:
		return None
	cl = AgglomerativeClustering(n_clusters=k, linkage="ward", distance_threshold=None)
	lab = cl.fit_predict(lab_u)
	centers = np.zeros((cl.n_clusters_, 3))
	for i in range(cl.n_clusters_):
		m = lab == i
		if m.any():
			centers[i] = np.average(uniq[m], weights=counts[m], axis=0)
	return np.clip(centers, 0, 255).astype(np.uint8)


def perceptual(rgba, num_colors=8, preserve_alpha=True, max_samples=10000):
	"""simplify_colors_perceptual (color_simplify.py:387-559), including its quirk: the nearest
	centre search compares LAB pixels with the RGB-valued uint8 centres (:540-544)."""
	_check(rgba)
	centers = perceptual_fit(rgba, num_colors, max_samples)
	if centers is None:
		return rgba, np.array([[0, 0, 0]])
	out, _ = assign_remap(rgba, centers, centers, "lab", preserve_alpha)
	return out, centers


def hsv_clustering(rgba, num_colors=8, preserve_alpha=True):
	"""simplify_colors_hsv_clustering (color_simplify.py:885-1036)."""
	from sklearn.cluster import KMeans

	_check(rgba)
	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	if not nt.any():
		return rgba, np.array([[0, 0, 0]])
	px = rgb[nt].reshape(-1, 3)
	hsv = ohsv.rgb_to_hsv_u8(px)
	keep = _dark_filter(hsv[:, 2], num_colors)
	feat = ohsv.hsv_weighted_features(hsv[keep])
	rgb_f = px[keep]
	k = min(num_colors, len(np.unique(feat, axis=0)))
	if k < 2:
		return rgba, np.array([[0, 0, 0]])
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		labels = KMeans(n_clusters=k, random_state=42, n_init=10).fit_predict(feat)
	centers = np.zeros((k, 3))
	for i in range(k):
		m = labels == i
		if m.any():
			centers[i] = np.mean(rgb_f[m], axis=0)
	centers = np.clip(centers, 0, 255).astype(np.uint8)
	all_lab = np.zeros(len(px), dtype=int)
	all_lab[np.where(keep)[0]] = labels
	dark = np.where(~keep)[0]
	if len(dark):
		all_lab[dark], _ = okm.argmin_min(px[dark], centers)
	q = np.zeros_like(rgb)
	q[nt] = centers[all_lab]
	return _epilogue(q, alpha, preserve_alpha), centers
