"""Batches of independent images: one k-means problem per image, all images of the batch in one kernel
launch per Lloyd iteration (cs_lloyd_iter_rgba8_batched; BASELINE config 4).  Images of a batch are
independent, so a multi-GPU job simply partitions the batch across ranks — no collective."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .engine import get_engine


def partition_batch(n_images: int, world: int, rank: int) -> tuple[int, int]:
	"""Contiguous slice [i0, i1) of the batch owned by `rank` (first n % world ranks get one more)."""
	base, extra = divmod(int(n_images), int(world))
	i0 = rank * base + min(rank, extra)
	return i0, i0 + base + (1 if rank < extra else 0)


def kmeans_rgb_batch(images, K: int, init_centers, n_iter: int, *, min_rgb_sum: int = -1, exact: bool = True,
                     device: int | None = None):
	"""RGB k-means (the Lloyd loop of simplify_colors_kmeans, reference color_simplify.py:79-80) on a batch.

	images        (B, H, W, 4) uint8 NumPy array or (B, n, 4) uint8 CUDA tensor (n % 4 == 0)
	init_centers  (B, K, 3) float64 initial centres (e.g. k-means++ seeds per image)
	Runs exactly n_iter iterations per image (no early stop: images converge at different times and the
	batch shares the launch), then one more E-step for the labels of the final centres.
	Returns (labels (B, n) uint8 tensor, centers (B, K, 3) float64 ndarray, counts (B, K) ndarray,
	n_empty (B,) ndarray of empty-cluster counts in the last iteration — no relocation on this path)."""
	import torch

	eng = get_engine(device)
	if isinstance(images, np.ndarray):
		b = images.shape[0]
		d = torch.from_numpy(np.ascontiguousarray(images).reshape(b, -1, 4)).to(eng.dev)
	else:
		d = images
	B, n = int(d.shape[0]), int(d.shape[1])
	if n % 4:
		raise ValueError("pixels per image must be a multiple of 4")
	K = int(K)
	c = [torch.from_numpy(np.ascontiguousarray(init_centers, dtype=np.float64).reshape(B, K, 3)).to(eng.dev),
	     torch.zeros((B, K, 3), dtype=torch.float64, device=eng.dev)]
	labels = torch.empty((B, n), dtype=torch.uint8, device=eng.dev)
	sums = torch.zeros((B, K, 3), dtype=torch.float64, device=eng.dev)
	counts = torch.zeros((B, K), dtype=torch.float64, device=eng.dev)
	stats = torch.zeros((B, 4), dtype=torch.float64, device=eng.dev)
	flags = _ffi.CS_LLOYD_EXACT_TIES if exact else 0
	cur = 0
	for it in range(n_iter + 1):
		last = it == n_iter
		eng._call("cs_lloyd_iter_rgba8_batched", d.data_ptr(), n, B, int(min_rgb_sum), c[cur].data_ptr(), K,
		          labels.data_ptr() if last else None, sums.data_ptr(), counts.data_ptr(), c[cur ^ 1].data_ptr(),
		          stats.data_ptr(), flags)
		if not last:
			cur ^= 1
			last_counts, last_stats = counts.clone(), stats.clone()
	if n_iter == 0:
		last_counts, last_stats = counts, stats
	return labels, c[cur].cpu().numpy(), last_counts.cpu().numpy(), last_stats[:, 1].cpu().numpy()
