"""Region analysis of a colour-simplified image on the CPU — restatement of analyze_regions
(app/processing/region_cleanup.py:9-130), the first consumer of the simplified image.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference loops over the unique colours of the opaque
pixels and calls cv.connectedComponentsWithStats on each colour's mask; OpenCV (4.13.0 here and on the GPU
box) is the routine the reference itself calls, so this restatement calls it the same way and is pinned by
tests/golden/reference_regions.npz (made from the unmodified reference, oracle/make_golden.py `regions`).
`component_order_keys` states the ordering rule of OpenCV's component numbers that the CUDA path relies on;
tests/test_oracle_regions.py checks it against cv2 itself.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np


def analyze_regions(rgba: np.ndarray, min_size_threshold: int = 100, connectivity: int = 8) -> dict:
	"""region_cleanup.py:9-130, same keys and value types."""
	import cv2 as cv

	if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4:
		raise ValueError("rgba must be HxWx4 uint8")
	empty = {"total_regions": 0, "small_regions": 0, "largest_region_size": 0, "smallest_region_size": 0,
	         "size_distribution": {}, "region_colors": [], "region_sizes": [], "all_regions": []}
	rgb, alpha = rgba[:, :, :3], rgba[:, :, 3]
	nt = alpha > 0
	if not np.any(nt):
		return empty
	all_regions, colors, sizes, small = [], [], [], 0
	for color in np.unique(rgb[nt].reshape(-1, 3), axis=0):
		mask = (np.all(rgb == color, axis=2) & nt).astype(np.uint8) * 255
		n, labels, stats, _ = cv.connectedComponentsWithStats(mask, connectivity=connectivity)
		for i in range(1, n):
			area = stats[i, cv.CC_STAT_AREA]
			if area > 0:
				all_regions.append({"color": tuple(color), "size": int(area), "label": i, "color_mask": mask, "labels": labels,
				                    "component_id": i,
				                    "bbox": (stats[i, cv.CC_STAT_LEFT], stats[i, cv.CC_STAT_TOP], stats[i, cv.CC_STAT_WIDTH],
				                             stats[i, cv.CC_STAT_HEIGHT])})
				colors.append(tuple(color))
				sizes.append(int(area))
				small += area < min_size_threshold
	if not sizes:
		return empty
	dist = defaultdict(int)
	for s in sizes:
		dist["< 50" if s < 50 else "50-99" if s < 100 else "100-199" if s < 200 else "200-499" if s < 500 else "500+"] += 1
	return {"total_regions": len(sizes), "small_regions": int(small), "largest_region_size": max(sizes),
	        "smallest_region_size": min(sizes), "size_distribution": dict(dist), "region_colors": colors,
	        "region_sizes": sizes, "all_regions": all_regions}


def component_order_keys(labels: np.ndarray, n_labels: int, connectivity: int) -> np.ndarray:
	"""The quantity by which OpenCV numbers components 1..n-1 of a mask: raster index of the component's
	first pixel (connectivity 4) or of its first 2x2 block (connectivity 8).  Strictly increasing in the
	component number."""
	h, w = labels.shape
	ys, xs = np.nonzero(labels)
	lab = labels[ys, xs]
	key = (ys // 2).astype(np.int64) * ((w + 1) // 2) + xs // 2 if connectivity == 8 else ys.astype(np.int64) * w + xs
	first = np.full(n_labels, np.iinfo(np.int64).max)
	np.minimum.at(first, lab, key)
	return first[1:]
