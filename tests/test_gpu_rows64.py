"""The fp64-row kernels behind simplify_colors_adaptive_distance (csrc/rows64.cu, engine.KMeansRows64) against
the library calls the reference makes (app/processing/color_simplify.py:809-814, 861-867)."""
import warnings

import numpy as np
import pytest

from gpu_util import engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,K,seed", [(6000, 8, 0), (4099, 5, 1), (20000, 16, 2)])
def test_kmeans_rows64_equals_sklearn_fit_predict(n, K, seed):
	"""Standardised blob data: the 10 seedings and the labels of the best run equal scikit-learn's."""
	import torch
	from sklearn.cluster import KMeans
	from sklearn.preprocessing import StandardScaler

	from image_segmenter_b200.engine import KMeansRows64

	rng = np.random.default_rng(seed)
	cent = rng.uniform(-60, 60, (K, 3))
	X = StandardScaler().fit_transform(cent[rng.integers(0, K, n)] + rng.normal(0, 4.0, (n, 3)))
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		ref = KMeans(n_clusters=K, random_state=42, n_init=10).fit(X)
	e = engine()
	got = KMeansRows64(e, torch.from_numpy(np.ascontiguousarray(X)).to(e.dev)).fit_predict(K)
	assert got.dtype == np.int64 and got.shape == (n,)
	mism = int((got != ref.labels_).sum())
	if mism:  # a tie at fp64 rounding level went the other way: same partition quality
		cen = np.array([X[got == k].mean(axis=0) for k in range(K)])
		inertia = float(((X - cen[got]) ** 2).sum())
		assert mism <= 3 and abs(inertia - ref.inertia_) <= 1e-9 * ref.inertia_
	else:
		assert np.array_equal(got, ref.labels_)


def test_nn_argmin_rows64_equals_sklearn():
	import torch
	from sklearn.metrics import pairwise_distances_argmin_min

	rng = np.random.default_rng(3)
	ref = rng.normal(0, 30, (5003, 3))
	ref[100:200] = ref[0:100]  # duplicate rows: the first copy wins
	q = np.vstack([rng.normal(0, 30, (777, 3)), ref[150:160], ref[5:9]])
	e = engine()
	got = e.nn_argmin_rows64(torch.from_numpy(q).to(e.dev), torch.from_numpy(ref).to(e.dev))
	exp, _ = pairwise_distances_argmin_min(q, ref)
	mism = np.nonzero(got != exp)[0]
	for i in mism:  # only exact / rounding-level ties between the GEMM form and the direct form
		a, b = ((q[i] - ref[got[i]]) ** 2).sum(), ((q[i] - ref[exp[i]]) ** 2).sum()
		assert abs(a - b) <= 1e-9 * max(1.0, b) and got[i] <= exp[i]
	assert np.array_equal(got[777:787], np.arange(50, 60))  # queries equal to duplicated rows -> the FIRST copy


@pytest.mark.parametrize("n,K,seed,max_iter", [(4672, 16, 0, 100), (5000, 8, 1, 100), (777, 5, 2, 300), (3000, 64, 3, 100), (40, 12, 4, 100)])
def test_sample_kmeans_equals_sklearn_fit(n, K, seed, max_iter):
	"""engine.SampleKMeans (cs_kmeans_fit_rows64_small: the ten Lloyd loops of the sample fit of
	simplify_colors_perceptual_fast in one launch) against KMeans(...).fit on CIELAB rows of distinct colours:
	centres to rounding, labels equal, same inertia, same number of iterations."""
	from sklearn.cluster import KMeans

	from image_segmenter_b200 import _colorspace as cspace
	from image_segmenter_b200.engine import SampleKMeans

	rng = np.random.default_rng(seed)
	X = cspace.rgb2lab_small(np.unique(rng.integers(0, 256, (n, 3), dtype=np.uint8), axis=0))
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		ref = KMeans(n_clusters=K, random_state=42, n_init=10, max_iter=max_iter).fit(X)
	got = SampleKMeans(engine()).fit(X, K, n_init=10, max_iter=max_iter, seed=42)
	assert got.cluster_centers_.shape == (K, 3) and got.labels_.shape == (len(X),)
	assert abs(got.inertia_ - ref.inertia_) <= 1e-9 * ref.inertia_
	assert np.allclose(got.cluster_centers_, ref.cluster_centers_, rtol=0, atol=1e-8)
	assert int((got.labels_ != ref.labels_).sum()) == 0
	assert got.n_iter_ == ref.n_iter_


def test_sample_kmeans_relocates_empty_clusters_like_sklearn():
	"""Few distinct rows, many clusters asked for: k-means++ picks duplicates of one row, clusters come out empty and
	_relocate_empty_clusters_dense moves the farthest rows into them — the in-kernel relocation must follow."""
	from sklearn.cluster import KMeans

	from image_segmenter_b200.engine import SampleKMeans

	rng = np.random.default_rng(5)
	base = rng.normal(0, 20, (9, 3))
	X = np.vstack([base[rng.integers(0, 9, 300)], rng.normal(0, 20, (6, 3))])
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		ref = KMeans(n_clusters=12, random_state=42, n_init=10, max_iter=100).fit(X)
	got = SampleKMeans(engine()).fit(X, 12, n_init=10, max_iter=100, seed=42)
	assert abs(got.inertia_ - ref.inertia_) <= 1e-9 * max(ref.inertia_, 1e-12)
	assert np.allclose(np.sort(got.cluster_centers_, axis=0), np.sort(ref.cluster_centers_, axis=0), atol=1e-8)
