import sys, warnings
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import color_simplify as cs
from oracle import pipeline as op, lab as olab, kmeans as okm
g = np.load(ROOT / "tests/golden/reference_entry_points.npz")
img = g["in_fewcolors"]
warnings.simplefilter("ignore")
np.random.seed(7)
out, pal = cs.simplify_colors_perceptual_fast(img, 6)
ref, rpal = g["fewcolors__perceptual_fast_6__rgba"], g["fewcolors__perceptual_fast_6__palette"]
print("pal equal", np.array_equal(pal, rpal)); print(pal); print(rpal)
bad = (out != ref).any(2)
print("bad px", bad.sum())
cols, cnt = np.unique(img[bad], axis=0, return_counts=True)
for c, n in zip(cols, cnt):
	m = (img == c).all(2) & bad
	print("src", c, n, "gpu->", np.unique(out[m], axis=0), "ref->", np.unique(ref[m], axis=0))
np.random.seed(7)
fit = op.perceptual_fast_fit(img, 6)
cl, crgb = fit
print("centres lab", cl)
for c in cols:
	x = olab.rgb2lab(c[:3].reshape(1, 3))
	d = ((x - cl) ** 2).sum(1)
	print(c, "dists", d, "argmin", d.argmin(), "argmin_min", okm.argmin_min(x, cl)[0])
