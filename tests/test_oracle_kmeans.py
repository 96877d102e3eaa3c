"""oracle/kmeans.py pinned against scikit-learn's own routines (the ones the reference reaches
from color_simplify.py:79-80, 544, 669-675, 692, 992-993) and against the committed fixture made
from sklearn's lloyd_iter_chunked_dense (tests/golden/sklearn_lloyd.npz)."""
import warnings

import numpy as np
import pytest

from oracle import kmeans as okm


def _data(seed, n=6000, K=9):
	rng = np.random.default_rng(seed)
	cent = rng.uniform(0, 255, (K, 3))
	X = np.clip(cent[rng.integers(0, K, n)] + rng.normal(0, 12, (n, 3)), 0, 255)
	C0 = X[rng.choice(n, K, replace=False)].copy()
	return X, C0


def test_single_step_matches_golden(golden_lloyd):
	g = golden_lloyd
	X = g["lab32"].astype(np.float64)
	labels, sums, counts, cnew, shift = okm.lloyd_iter(X, g["C0"])
	assert np.array_equal(labels, g["step_labels"])
	assert np.array_equal(counts, g["step_weights"])
	assert np.allclose(cnew, g["step_centers"], rtol=1e-12, atol=1e-12)
	assert abs(shift - float((g["step_shift"] ** 2).sum())) <= 1e-9 * max(1.0, shift)


def test_full_fit_matches_golden(golden_lloyd):
	g = golden_lloyd
	X = g["lab32"].astype(np.float64)
	labels, inertia, centers, n_iter = okm.kmeans_single_lloyd(X, g["C0"], max_iter=300, tol=okm.sklearn_tol(X))
	assert n_iter == int(g["fit_n_iter"])
	assert np.array_equal(labels, g["fit_labels"])
	assert np.allclose(centers, g["fit_centers"], rtol=1e-10, atol=1e-10)
	assert abs(inertia - float(g["fit_inertia"])) <= 1e-9 * inertia


@pytest.mark.parametrize("seed", range(3))
def test_step_matches_sklearn_live(seed):
	from sklearn.cluster._k_means_lloyd import lloyd_iter_chunked_dense

	X, C0 = _data(seed)
	K = C0.shape[0]
	cnew, w = np.zeros_like(C0), np.zeros(K)
	lab, shift = np.full(len(X), -1, np.int32), np.zeros(K)
	lloyd_iter_chunked_dense(X, np.ones(len(X)), C0, cnew, w, lab, shift, 2)
	labels, _, counts, centers, _ = okm.lloyd_iter(X, C0)
	assert np.array_equal(labels, lab) and np.array_equal(counts, w)
	assert np.allclose(centers, cnew, rtol=1e-12, atol=1e-12)


def test_fit_matches_sklearn_kmeans_live():
	from sklearn.cluster import KMeans

	X, C0 = _data(5)
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		km = KMeans(n_clusters=C0.shape[0], init=C0, n_init=1, max_iter=300, tol=1e-4).fit(X)
	labels, inertia, centers, n_iter = okm.kmeans_single_lloyd(X, C0, 300, okm.sklearn_tol(X))
	assert n_iter == km.n_iter_
	assert np.array_equal(labels, km.labels_)
	assert np.allclose(centers, km.cluster_centers_, rtol=1e-9, atol=1e-9)


def test_empty_cluster_average_and_relocation():
	X = np.array([[0.0, 0, 0], [1, 0, 0], [10, 0, 0], [11, 0, 0]])
	C0 = np.array([[0.5, 0, 0], [10.5, 0, 0], [1000.0, 0, 0]])  # third centre attracts nothing
	labels, sums, counts, cnew, _ = okm.lloyd_iter(X, C0, relocate=True)
	assert counts.sum() == 4 and (counts > 0).all()  # relocation filled the empty cluster
	from sklearn.cluster._k_means_lloyd import lloyd_iter_chunked_dense

	c2, w = np.zeros_like(C0), np.zeros(3)
	lab, sh = np.full(4, -1, np.int32), np.zeros(3)
	lloyd_iter_chunked_dense(X, np.ones(4), C0, c2, w, lab, sh, 1)
	assert np.array_equal(np.sort(w), np.sort(counts))
	# _average_centers without relocation: empty cluster copies the heaviest one as visited
	_, s_nr, c_nr, cn_nr, _ = okm.lloyd_iter(X, C0, relocate=False)
	assert c_nr[2] == 0 and np.array_equal(cn_nr[2], cn_nr[int(np.argmax(c_nr))])


def test_argmin_min_matches_sklearn():
	from sklearn.metrics import pairwise_distances_argmin_min

	rng = np.random.default_rng(4)
	X = rng.integers(0, 256, (5000, 3)).astype(np.float64)
	Y = rng.integers(0, 256, (13, 3)).astype(np.float64)
	Y[7] = Y[2]  # duplicate palette entry: lowest index must win
	X[:50] = Y[7]
	i_ref, d_ref = pairwise_distances_argmin_min(X, Y)
	i, d = okm.argmin_min(X, Y)
	assert np.array_equal(i, i_ref)
	assert np.allclose(d, d_ref, rtol=1e-9, atol=1e-6)
	assert (i[:50] == 2).all()
