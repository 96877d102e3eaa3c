"""Lloyd k-means and nearest-centre search in NumPy float64 — restatement of the scikit-learn
routines the reference reaches from app/processing/color_simplify.py:79-80, 669-675, 811-812,
992-993 (KMeans) and :544, 692, 866, 1020, 1107 (pairwise_distances_argmin_min).

TEST INFRASTRUCTURE (see oracle/__init__.py).  scikit-learn is an un-vendored, unpinned
dependency of the reference (requirements.txt: `scikit-learn>=1.3`); the image has 1.9.0, whose
Cython sources ship in the wheel and are cited below.  Pinned by tests/test_oracle_kmeans.py,
which runs these functions against sklearn's own `lloyd_iter_chunked_dense`, `KMeans` and
`pairwise_distances_argmin_min` on the same inputs.
"""
from __future__ import annotations

import numpy as np

CHUNK = 256  # sklearn/cluster/_k_means_common.pyx:13 (CHUNK_SIZE)


def assign_labels(X: np.ndarray, centers: np.ndarray) -> np.ndarray:
	"""E-step of _update_chunk_dense (sklearn/cluster/_k_means_lloyd.pyx:166-213): per chunk of
	256 samples d_ij = |c_j|^2 - 2 x_i.c_j (the |x_i|^2 term is dropped), label = first j with the
	strictly smallest value."""
	X = np.ascontiguousarray(X, dtype=np.float64)
	C = np.ascontiguousarray(centers, dtype=np.float64)
	cn = (C * C).sum(axis=1)
	labels = np.empty(X.shape[0], dtype=np.int32)
	step = CHUNK * 256  # many sklearn chunks per NumPy call; chunking does not change a row's result
	for s in range(0, X.shape[0], step):
		d = cn[None, :] - 2.0 * (X[s:s + step] @ C.T)
		labels[s:s + step] = np.argmin(d, axis=1)  # np.argmin returns the first minimum
	return labels


def accumulate(X: np.ndarray, labels: np.ndarray, K: int):
	"""M-step accumulation of _update_chunk_dense (:215-218) with unit sample weights:
	weight_in_clusters[label] += 1; centers_new[label] += x."""
	X = np.asarray(X, dtype=np.float64)
	counts = np.bincount(labels, minlength=K).astype(np.float64)
	sums = np.stack([np.bincount(labels, weights=X[:, j], minlength=K) for j in range(X.shape[1])], axis=1)
	return sums, counts


def relocate_empty(X, centers_old, sums, counts, labels):
	"""_relocate_empty_clusters_dense (sklearn/cluster/_k_means_common.pyx:167-211), unit weights.
	Deviation: sklearn takes the n_empty farthest samples in np.argpartition's (unspecified)
	order; this restatement — like the CUDA kernel — takes them in descending distance, lowest
	index first on equal distances.  Identical when one cluster is empty and the maximum is unique."""
	empty = np.where(counts == 0)[0]
	if len(empty) == 0:
		return sums, counts
	X = np.asarray(X, dtype=np.float64)
	dist = ((X - centers_old[labels]) ** 2).sum(axis=1)
	if dist.max() == 0:
		return sums, counts
	order = np.lexsort((np.arange(len(dist)), -dist))[:len(empty)]
	sums = sums.copy()
	counts = counts.copy()
	for new_id, far in zip(empty, order):
		old_id = labels[far]
		sums[old_id] -= X[far]
		sums[new_id] = X[far]
		counts[new_id] = 1.0
		counts[old_id] -= 1.0
	return sums, counts


def average_centers(sums, counts):
	"""_average_centers (_k_means_common.pyx:274-295): centre = sum * (1/w); an empty cluster
	copies centers[argmax_weight] AS IT IS WHEN VISITED (raw sum for j < argmax, averaged after)."""
	C = np.array(sums, dtype=np.float64, copy=True)
	am = int(np.argmax(counts))
	for j in range(C.shape[0]):
		if counts[j] > 0:
			C[j] *= 1.0 / counts[j]
		else:
			C[j] = C[am]
	return C


def center_shift_total(c_old, c_new) -> float:
	"""_center_shift (:298-311) then (center_shift**2).sum() of _kmeans_single_lloyd
	(sklearn/cluster/_kmeans.py:731)."""
	shift = np.empty(c_old.shape[0])
	for j in range(c_old.shape[0]):
		s = 0.0
		for k in range(c_old.shape[1]):
			t = c_new[j, k] - c_old[j, k]
			s += t * t
		shift[j] = np.sqrt(s)
	return float((shift ** 2).sum())


def lloyd_iter(X, centers_old, relocate=True):
	"""One lloyd_iter_chunked_dense call (sklearn/cluster/_k_means_lloyd.pyx:23-153).
	Returns labels, sums, counts (after relocation), centers_new, shift_total."""
	K = centers_old.shape[0]
	labels = assign_labels(X, centers_old)
	sums, counts = accumulate(X, labels, K)
	if relocate:
		sums, counts = relocate_empty(X, centers_old, sums, counts, labels)
	centers_new = average_centers(sums, counts)
	return labels, sums, counts, centers_new, center_shift_total(centers_old, centers_new)


def inertia(X, centers, labels) -> float:
	"""_inertia_dense (_k_means_common.pyx:94-124): sum of squared distances to assigned centre."""
	X = np.asarray(X, dtype=np.float64)
	return float(((X - centers[labels]) ** 2).sum())


def kmeans_single_lloyd(X, centers_init, max_iter=300, tol=0.0):
	"""_kmeans_single_lloyd (sklearn/cluster/_kmeans.py:630-758) from given centres, already
	mean-centred by the caller or not (Lloyd is translation-equivariant up to rounding).
	Returns labels, inertia, centers, n_iter."""
	X = np.ascontiguousarray(X, dtype=np.float64)
	centers = np.array(centers_init, dtype=np.float64, copy=True)
	labels_old = np.full(X.shape[0], -1, dtype=np.int32)
	strict = False
	i = -1
	for i in range(max_iter):
		labels, _, _, centers_new, shift_tot = lloyd_iter(X, centers)
		centers = centers_new
		if np.array_equal(labels, labels_old):
			strict = True
			break
		if shift_tot <= tol:
			break
		labels_old = labels
	if not strict:
		labels = assign_labels(X, centers)
	return labels, inertia(X, centers, labels), centers, i + 1


def sklearn_tol(X, tol=1e-4) -> float:
	"""_tolerance (sklearn/cluster/_kmeans.py:285-293): mean of per-feature variances * tol."""
	return float(np.mean(np.var(np.asarray(X, dtype=np.float64), axis=0)) * tol)


def argmin_min(X, Y):
	"""pairwise_distances_argmin_min, euclidean (sklearn/metrics/pairwise.py:711-845 -> ArgKmin,
	_argkmin.pyx.tp:471-510): both operands float64, d2 = |x|^2 - 2 x.y + |y|^2 clamped at 0,
	k=1 heap rejects `val >= current` (sklearn/utils/_heap.pyx:46-47) => lowest index on ties.
	Returns (indices intp, distances float64)."""
	X = np.ascontiguousarray(X, dtype=np.float64)
	Y = np.ascontiguousarray(Y, dtype=np.float64)
	yn = (Y * Y).sum(axis=1)
	idx = np.empty(X.shape[0], dtype=np.intp)
	dist = np.empty(X.shape[0], dtype=np.float64)
	step = 1 << 16
	for s in range(0, X.shape[0], step):
		xs = X[s:s + step]
		d2 = (xs * xs).sum(axis=1)[:, None] - 2.0 * (xs @ Y.T) + yn[None, :]
		np.maximum(d2, 0.0, out=d2)
		j = np.argmin(d2, axis=1)
		idx[s:s + step] = j
		dist[s:s + step] = np.sqrt(d2[np.arange(len(j)), j])
	return idx, dist


def near_tie_gap(X, centers):
	"""Helper for parity tests: (best, second) fp64 direct squared distances per sample, so a
	label mismatch can be checked against the documented near-tie bound."""
	X = np.asarray(X, dtype=np.float64)
	C = np.asarray(centers, dtype=np.float64)
	d = np.zeros((X.shape[0], C.shape[0]))
	for j in range(X.shape[1]):
		d += (X[:, j, None] - C[None, :, j]) ** 2
	if C.shape[0] == 1:
		return d[:, 0], np.full(X.shape[0], np.inf)
	part = np.partition(d, 1, axis=1)
	return part[:, 0], part[:, 1]
