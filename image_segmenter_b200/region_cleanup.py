"""Drop-in for `analyze_regions` of `app/processing/region_cleanup.py` (reference :9-130) — the first
consumer of the colour-simplified image and SURVEY.md §8f rank 4.

The reference runs cv.connectedComponentsWithStats once per unique colour (O(K * N)); here ONE device
labelling (csrc/ccl.cu) finds the components of all colours, a second pass gives areas / bounding boxes /
the key that reproduces OpenCV's component numbering, and the per-colour `labels` / `color_mask` arrays the
reference returns are cut out on the device per colour.  Same keys, value types and ordering as the
reference: regions are listed colour by colour (lexicographic RGB, np.unique order), inside a colour in
OpenCV's component order.  No CPU fallback.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

from .engine import get_engine

__all__ = ["analyze_regions"]


def analyze_regions(rgba: np.ndarray, min_size_threshold: int = 100, connectivity: int = 8) -> dict:
	import torch

	if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4:
		raise ValueError("rgba must be HxWx4 uint8")
	if connectivity not in (4, 8):
		raise ValueError("connectivity must be 4 or 8")  # cv2 raises its own error for anything else
	empty = {"total_regions": 0, "small_regions": 0, "largest_region_size": 0, "smallest_region_size": 0,
	         "size_distribution": {}, "region_colors": [], "region_sizes": [], "all_regions": []}
	h, w = rgba.shape[:2]
	n = h * w
	eng = get_engine()
	d = eng.upload_rgba(rgba)
	labels = torch.empty(n, dtype=torch.int32, device=eng.dev)
	eng._call("cs_ccl_label", d.data_ptr(), w, h, int(connectivity), labels.data_ptr())
	cnt = torch.zeros(1, dtype=torch.int64, device=eng.dev)
	rank = torch.empty(n, dtype=torch.int32, device=eng.dev)
	eng._call("cs_ccl_roots", labels.data_ptr(), n, None, None, 0, cnt.data_ptr())
	R = int(cnt.item())
	if R == 0:
		return empty
	roots = torch.empty(R, dtype=torch.int32, device=eng.dev)
	eng._call("cs_ccl_roots", labels.data_ptr(), n, rank.data_ptr(), roots.data_ptr(), R, cnt.data_ptr())
	area = torch.empty(R, dtype=torch.int32, device=eng.dev)
	bbox = torch.empty((R, 4), dtype=torch.int32, device=eng.dev)
	okey = torch.empty(R, dtype=torch.int64, device=eng.dev)
	eng._call("cs_ccl_stats", labels.data_ptr(), rank.data_ptr(), w, h, int(connectivity), R, area.data_ptr(), bbox.data_ptr(),
	          okey.data_ptr())
	# component colours = colour of the root pixel (palette-sized host work from here on: R components)
	root_px = eng.gather(d, roots.cpu().numpy().astype(np.int64)).cpu().numpy()[:, :3]
	h_area, h_bbox, h_key = area.cpu().numpy().astype(np.int64), bbox.cpu().numpy(), okey.cpu().numpy()
	ckey = (root_px[:, 0].astype(np.int64) << 16) | (root_px[:, 1].astype(np.int64) << 8) | root_px[:, 2].astype(np.int64)
	ukeys, comp_color = np.unique(ckey, return_inverse=True)  # ascending key == np.unique(rows) order
	order = np.lexsort((h_key, comp_color))  # colour by colour, OpenCV's component order inside a colour
	comp_local = np.empty(R, dtype=np.int32)
	start = np.searchsorted(comp_color[order], np.arange(len(ukeys)))
	comp_local[order] = (np.arange(R) - start[comp_color[order]] + 1).astype(np.int32)
	d_cc = torch.from_numpy(comp_color.astype(np.int32)).to(eng.dev)
	d_cl = torch.from_numpy(comp_local).to(eng.dev)

	per_colour = {}

	def arrays_of(c: int):
		if c not in per_colour:
			lab_c = torch.empty(n, dtype=torch.int32, device=eng.dev)
			mask_c = torch.empty(n, dtype=torch.uint8, device=eng.dev)
			eng._call("cs_ccl_extract", labels.data_ptr(), rank.data_ptr(), n, d_cc.data_ptr(), d_cl.data_ptr(), int(c),
			          lab_c.data_ptr(), mask_c.data_ptr())
			# (pageable on purpose: up to 256 colours x 5 B/px stay alive in the result — too much to page-lock)
			per_colour[c] = (mask_c.cpu().numpy().reshape(h, w), lab_c.cpu().numpy().reshape(h, w))
		return per_colour[c]

	all_regions, colors, sizes, small = [], [], [], 0
	for r in order:
		c = int(comp_color[r])
		k = int(ukeys[c])
		color = (np.uint8(k >> 16), np.uint8((k >> 8) & 0xFF), np.uint8(k & 0xFF))
		mask_c, lab_c = arrays_of(c)
		a = int(h_area[r])
		i = int(comp_local[r])
		x0, y0, x1, y1 = (int(v) for v in h_bbox[r])
		all_regions.append({"color": color, "size": a, "label": i, "color_mask": mask_c, "labels": lab_c, "component_id": i,
		                    "bbox": (np.int32(x0), np.int32(y0), np.int32(x1 - x0 + 1), np.int32(y1 - y0 + 1))})
		colors.append(color)
		sizes.append(a)
		small += a < min_size_threshold
	dist = defaultdict(int)
	for s in sizes:
		dist["< 50" if s < 50 else "50-99" if s < 100 else "100-199" if s < 200 else "200-499" if s < 500 else "500+"] += 1
	return {"total_regions": len(sizes), "small_regions": int(small), "largest_region_size": max(sizes),
	        "smallest_region_size": min(sizes), "size_distribution": dict(dist), "region_colors": colors,
	        "region_sizes": sizes, "all_regions": all_regions}
