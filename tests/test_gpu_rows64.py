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
