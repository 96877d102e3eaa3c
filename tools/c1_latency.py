"""BASELINE config 1 on its real input: wall time of simplify_colors_kmeans(working_image_cleaned, 16) and of the
other entry points whose answers the reference gave (host array in, host array out)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import color_simplify as cs

g = np.load(ROOT / "tests" / "golden" / "working_image_cleaned.npz")
rgb = g["colours"][g["index"]]
img = np.ascontiguousarray(np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)]))
out = {}
for name, fn in (("kmeans_16", lambda: cs.simplify_colors_kmeans(img, 16)), ("hsv_clustering_16", lambda: cs.simplify_colors_hsv_clustering(img, 16)),
                 ("median_cut_16", lambda: cs.simplify_colors_median_cut(img, 16)), ("threshold_16", lambda: cs.simplify_colors_threshold(img, 16)),
                 ("statistics", lambda: cs.get_color_statistics(img)), ("adaptive", lambda: cs.simplify_colors_adaptive(img, 16, True, "adaptive"))):
	fn()
	best = 1e9
	for _ in range(5):
		torch.cuda.synchronize()
		t0 = time.perf_counter()
		fn()
		torch.cuda.synchronize()
		best = min(best, time.perf_counter() - t0)
	out[name + "_ms"] = round(best * 1e3, 3)
print(json.dumps(out))
