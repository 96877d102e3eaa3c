"""Host-side logic of the drop-in module that needs no device: the filter cascade, the exact
moments, the k-means++ stream replay, the trivial helpers."""
import warnings

import numpy as np
import pytest

from image_segmenter_b200 import color_simplify as cs


def test_brightness_threshold_cascade():
	# enough bright pixels -> > hi ; too few -> > lo ; none -> everything
	assert cs._brightness_threshold(100, 200, 8, 90, 30) == 90
	assert cs._brightness_threshold(3, 200, 8, 90, 30) == 30
	assert cs._brightness_threshold(0, 0, 8, 90, 30) == -1
	assert cs._brightness_threshold(3, 0, 8, 90, 30) == -1
	# reference semantics on a real array (color_simplify.py:56-64)
	rng = np.random.default_rng(0)
	for _ in range(20):
		rgb = rng.integers(0, 60, (50, 3), dtype=np.uint8)
		k = int(rng.integers(2, 60))
		b = np.mean(rgb, axis=1)
		m = b > 30
		if m.sum() < k:
			m = b > 10
		if m.sum() == 0:
			m = np.ones(len(b), bool)
		s = rgb.astype(int).sum(1)
		thr = cs._brightness_threshold(int((s > 90).sum()), int((s > 30).sum()), k, 90, 30)
		assert np.array_equal(m, s > thr)


def test_moments_from_hist_match_numpy():
	rng = np.random.default_rng(1)
	px = rng.integers(0, 256, (20000, 3), dtype=np.uint8)
	hist = np.stack([np.bincount(px[:, c], minlength=256) for c in range(3)])
	ident = np.tile(np.arange(256, dtype=np.float64), (3, 1))
	mean, var = cs._moments_from_hist(hist, ident)
	X = px.astype(np.float64)
	assert np.allclose(mean, X.mean(0), rtol=1e-14) and np.allclose(var, X.var(0), rtol=1e-12)
	from image_segmenter_b200 import _colorspace as csp

	lut = csp.hsv_feature_luts().astype(np.float64)
	F = np.stack([lut[c][px[:, c]] for c in range(3)], 1)
	mean, var = cs._moments_from_hist(hist, lut)
	assert np.allclose(mean, F.mean(0), rtol=1e-13) and np.allclose(var, F.var(0), rtol=1e-11)


def test_kmeanspp_replay_reproduces_sklearn_fit():
	"""Seeds drawn back to back from RandomState(42) + sklearn's own single-run Lloyd == KMeans.fit."""
	from sklearn.cluster import KMeans
	from sklearn.cluster._kmeans import _kmeans_single_lloyd, _tolerance
	from sklearn.cluster._k_means_common import _is_same_clustering

	rng = np.random.default_rng(3)
	X = rng.integers(0, 256, (3000, 3)).astype(np.float64)
	K = 6
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		km = KMeans(n_clusters=K, random_state=42, n_init=10).fit(X)
	mean = X.mean(axis=0)
	Xc = X - mean
	best = None
	for idx in cs._seed_kmeans_plusplus(X, K, 10):
		lab, inertia, cen, _ = _kmeans_single_lloyd(Xc, np.ones(len(X)), Xc[idx].copy(), max_iter=300,
		                                             tol=_tolerance(Xc, 1e-4), n_threads=1)
		if best is None or (inertia < best[1] and not _is_same_clustering(lab, best[0], K)):
			best = (lab, inertia, cen)
	assert np.array_equal(best[0], km.labels_)
	assert np.allclose(best[2] + mean, km.cluster_centers_, rtol=1e-12, atol=1e-10)


def test_small_helpers():
	pal = cs.create_palette_from_colors([(1, 2, 3), (250, 251, 252)])
	assert pal.dtype == np.uint8 and pal.shape == (2, 3)
	assert cs.get_recommended_algorithm((2000, 2000)) == "perceptual_fast"
	assert cs.get_recommended_algorithm((2000, 2000), gpu_available=True) == "perceptual"
	assert cs.get_recommended_algorithm((800, 800)) == "perceptual_fast"
	assert cs.get_recommended_algorithm((400, 400)) == "hsv_clustering"
	assert cs.get_recommended_algorithm((100, 100)) == "kmeans"
	info = cs.check_gpu_availability()
	assert set(info) == {"cupy_available", "pytorch_available", "cuda_available", "gpu_count", "gpu_names"}
	assert np.array_equal(cs._truncate_u8(np.array([[153.9999, -3.0, 300.0]])), [[153, 0, 255]])


def test_four_channel_area_resize_equals_the_reference_two_calls():
	"""simplify_colors_perceptual_fast downsamples with ONE 4-channel cv.resize(INTER_AREA) instead of the reference's
	cv.resize(rgb) + cv.resize(alpha) (color_simplify.py:608-614): the planes must be bit-identical."""
	import cv2 as cv

	rng = np.random.default_rng(0)
	for h, w in ((2160, 3840), (1000, 1777), (513, 700), (600, 5000), (1023, 511)):
		img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
		img[::7, :, 3] = 0
		scale = min(512 / h, 512 / w)
		nh, nw = int(h * scale), int(w * scale)
		both = cv.resize(img, (nw, nh), interpolation=cv.INTER_AREA)
		assert np.array_equal(both[:, :, :3], cv.resize(img[:, :, :3], (nw, nh), interpolation=cv.INTER_AREA))
		assert np.array_equal(both[:, :, 3], cv.resize(img[:, :, 3], (nw, nh), interpolation=cv.INTER_AREA))


def test_same_clustering_matches_sklearn():
	"""engine._same_clustering_np (the best-of-n_init rule of SampleKMeans) against scikit-learn's own
	_is_same_clustering (sklearn/cluster/_k_means_common.pyx:314-328), including its one-way nature."""
	from sklearn.cluster._k_means_common import _is_same_clustering

	from image_segmenter_b200.engine import _same_clustering_np

	rng = np.random.default_rng(0)
	for trial in range(200):
		K = int(rng.integers(2, 9))
		a = rng.integers(0, K, 60).astype(np.int32)
		kind = trial % 4
		if kind == 0:
			b = rng.permutation(K).astype(np.int32)[a]            # a relabelling: same clustering
		elif kind == 1:
			b = rng.integers(0, K, 60).astype(np.int32)            # unrelated
		elif kind == 2:
			b = (a // 2).astype(np.int32)                          # b merges clusters of a: a -> b is still a function
		else:
			b = a.copy()
			b[int(rng.integers(0, 60))] = (b[0] + 1) % K           # one point moved
		assert _same_clustering_np(a, b, K) == bool(_is_same_clustering(a, b, K)), (trial, a, b)
