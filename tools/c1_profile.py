"""cProfile of simplify_colors_kmeans on the config-1 image (where do the host milliseconds go?)."""
import cProfile
import pstats
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import color_simplify as cs

g = np.load(ROOT / "tests" / "golden" / "working_image_cleaned.npz")
rgb = g["colours"][g["index"]]
img = np.ascontiguousarray(np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)]))
cs.simplify_colors_kmeans(img, 16)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
	cs.simplify_colors_kmeans(img, 16)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
