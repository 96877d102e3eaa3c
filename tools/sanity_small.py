"""Exercise every kernel of libcolorsimplify on small inputs (for compute-sanitizer runs)."""
import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
warnings.simplefilter("ignore")
from gpu_util import blobby_rgba, lab_like, lloyd_step, planes_of  # noqa: E402
from image_segmenter_b200 import color_simplify as cs  # noqa: E402
from image_segmenter_b200 import region_cleanup as rc  # noqa: E402
from image_segmenter_b200.batch import kmeans_rgb_batch  # noqa: E402

img = blobby_rgba(1, 61, 47)
np.random.seed(0)
for fn in (cs.simplify_colors_kmeans, cs.simplify_colors_median_cut, cs.simplify_colors_octree, cs.simplify_colors_threshold,
           cs.simplify_colors_perceptual, cs.simplify_colors_perceptual_fast, cs.simplify_colors_hsv_clustering):
	out, pal = fn(img, 6)
	print(fn.__name__, out.shape, len(pal))
cp = np.array([[250, 10, 10], [10, 240, 30], [20, 30, 230], [128, 128, 128]], np.uint8)
for m in ("lab", "hsv", "rgb"):
	cs.simplify_colors_custom_palette(img, cp, True, m)
print(cs.get_color_statistics(img)["total_unique_colors"])
rng = np.random.default_rng(0)
for n, K in ((5003, 16), (4099, 64), (2050, 256), (3, 5)):
	X = lab_like(rng, n)
	C = X[rng.choice(n, K, replace=n < K)].astype(np.float64) + (rng.normal(0, 1e-3, (K, 3)) if n < K else 0)
	for exact in (False, True):
		lloyd_step(planes_of(X), n, C, exact=exact, inertia=True)
		lloyd_step(planes_of(X), n, C, exact=exact, fused=True)
imgs = rng.integers(0, 256, (3, 40, 52, 4), dtype=np.uint8)
imgs[..., 3] = 255
kmeans_rgb_batch(imgs, 5, imgs[:, 0, :5, :3].astype(np.float64) + 0.1, 3)
q = img.copy()
q[:, :, :3] = (q[:, :, :3] // 85) * 85
print(rc.analyze_regions(q)["total_regions"], rc.analyze_regions(q, 100, 4)["total_regions"])
print("sanity ok")
