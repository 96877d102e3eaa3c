"""The reference's CPU implementation of the Lloyd iteration, timed as the CPU baseline.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py): only bench.py's `cpu_baseline` leg
and `bench.py --impl reference` call this.  The reference's arithmetic for this path lives in
scikit-learn (KMeans.fit at app/processing/color_simplify.py:79-80, 669-675, 811-812, 992-993 ->
sklearn/cluster/_k_means_lloyd.pyx:23-218 `lloyd_iter_chunked_dense`), which is installed in the
image (1.9.0) — so the baseline runs THAT routine, fp64 as the reference runs it, with all the
OpenMP threads it will use; the NumPy port (oracle/kmeans.py) is the fallback if the private
entry point is ever missing.
"""
from __future__ import annotations

import os
import time

import numpy as np


def host_threads() -> int:
	"""Cores this process may run on.  (NOT sklearn's _openmp_effective_n_threads(): that obeys
	OMP_NUM_THREADS, which torchrun sets to 1 in every worker it spawns — VERDICT r1 weak #3.)"""
	try:
		return len(os.sched_getaffinity(0))
	except Exception:
		return os.cpu_count() or 1


def make_lab_sample(n: int, seed: int) -> np.ndarray:
	"""fp64 CIELAB rows of `n` seeded uniform-random sRGB pixels (SURVEY.md §8d synthetic inputs)."""
	from . import lab as olab

	rng = np.random.default_rng(seed)
	out = np.empty((n, 3), dtype=np.float64)
	step = 1 << 20
	for s in range(0, n, step):
		m = min(step, n - s)
		out[s:s + m] = olab.rgb2lab(rng.integers(0, 256, (m, 3), dtype=np.uint8))
	return out


def time_lloyd_iterations(X: np.ndarray, C0: np.ndarray, iters: int, threads: int | None = None):
	"""Runs `iters` Lloyd iterations (assign + update, labels written) from C0 on the host with `threads`
	OpenMP threads (default: every core of the process; the prange of lloyd_iter_chunked_dense takes the
	count as an explicit num_threads clause, so OMP_NUM_THREADS does not cap it) and BLAS limited to one
	thread per chunk, as KMeans.fit runs it (sklearn/cluster/_kmeans.py:697 threadpool_limits).
	Returns (seconds, kind, threads, final centres)."""
	K = C0.shape[0]
	n = X.shape[0]
	try:
		from sklearn.cluster._k_means_lloyd import lloyd_iter_chunked_dense
		from threadpoolctl import threadpool_limits

		threads = int(threads) if threads else host_threads()
		w = np.ones(n, dtype=np.float64)
		c_old, c_new = np.array(C0, dtype=np.float64, copy=True), np.zeros_like(C0, dtype=np.float64)
		wk, labels, shift = np.zeros(K), np.full(n, -1, np.int32), np.zeros(K)
		with threadpool_limits(limits=1, user_api="blas"):
			lloyd_iter_chunked_dense(X, w, c_old, c_new, wk, labels, shift, threads)  # warm-up (thread pool, pages)
			c_old[:] = C0
			t0 = time.perf_counter()
			for _ in range(iters):
				lloyd_iter_chunked_dense(X, w, c_old, c_new, wk, labels, shift, threads)
				c_old, c_new = c_new, c_old
			dt = time.perf_counter() - t0
		return dt, "reference", threads, c_old
	except ImportError:
		from . import kmeans as okm

		c = np.array(C0, dtype=np.float64, copy=True)
		t0 = time.perf_counter()
		for _ in range(iters):
			_, _, _, c, _ = okm.lloyd_iter(X, c)
		return time.perf_counter() - t0, "port", 1, c
