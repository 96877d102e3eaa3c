"""Per-init comparison of the GPU RGB k-means with the oracle on a golden image (development aid)."""
import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import color_simplify as cs
from image_segmenter_b200.engine import KMeansGPU, get_engine
from oracle import kmeans as okm

g = np.load(ROOT / "tests/golden/reference_entry_points.npz")
name, k = sys.argv[1] if len(sys.argv) > 1 else "blobby", int(sys.argv[2]) if len(sys.argv) > 2 else 5
img = g[f"in_{name}"]
eng = get_engine()
d = eng.upload_rgba(img)
n_op, n_hi, n_lo, _ = eng.mask_stats(d, -1)
thr = cs._brightness_threshold(n_hi, n_lo, k, 90, 30)
px, _ = eng.select_compact(d, 0, thr)
X = px.cpu().numpy()[:, :3].astype(np.float64)
K = min(k, eng.mask_stats(d, thr, want_unique=True)[3])
ident = np.tile(np.arange(256, dtype=np.float64), (3, 1))
_, var = cs._moments_from_hist(eng.channel_hist(d, 0, thr), ident)
tol = float(np.mean(var) * 1e-4)
print("thr", thr, "K", K, "n", len(X), "tol", tol, "sk tol", okm.sklearn_tol(X))
km = KMeansGPU(eng, "rgba8", d.shape[0], px=d, mask_mode=0, min_bright=thr)
mean = X.mean(0)
for i, idx in enumerate(cs._seed_kmeans_plusplus(X, K, 10)):
	fit = km.fit_single(X[idx], 300, tol)
	lab, inert, cen, nit = okm.kmeans_single_lloyd(X - mean, X[idx] - mean, 300, tol)
	cen = cen + mean
	gl = fit.labels[:d.shape[0]].cpu().numpy()
	gl = gl[gl != 255]
	print(i, "gpu n_iter", fit.n_iter, "inertia", fit.inertia, "| oracle", nit, inert, "| max centre diff",
	      np.abs(fit.centers - cen).max(), "label diff", int((gl != lab).sum()))
	# trajectory: first diverging iteration
	c_g, c_o = X[idx].copy(), X[idx].copy()
	for it in range(1, 8):
		f1 = km.fit_single(c_g, 1, -1.0)
		_, _, _, c_o2, _ = okm.lloyd_iter(X, c_o)
		print("   it", it, "centre diff", np.abs(f1.centers - c_o2).max())
		c_g, c_o = f1.centers, c_o2
	if i >= 2:
		break
