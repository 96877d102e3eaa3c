import sys, warnings
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200.engine import get_engine, KMeansRows64
from sklearn.cluster import DBSCAN, KMeans
from sklearn.preprocessing import StandardScaler
warnings.simplefilter("ignore")
g = np.load(ROOT / "tests/golden/reference_entry_points.npz")
eng = get_engine(0)
for name, k in (("blobby", 8), ("fewcolors", 8)):
	img = g[f"in_{name}"]
	d = eng.upload_rgba(img)
	src = eng.select_compact(d, 0, -1)[0] if (img[..., 3] == 0).any() else d
	lab = eng.rgba_to_lab_f64(src).cpu().numpy()
	keep = lab[:, 0] > 10
	lab_f = lab[keep]
	lab_n = StandardScaler().fit_transform(lab_f)
	cl = DBSCAN(eps=0.125, min_samples=3).fit_predict(lab_n)
	print(name, k, "n", len(lab_n), "distinct rows", len(np.unique(lab_n, axis=0)), "dbscan clusters", len(np.unique(cl)))
	ref = KMeans(n_clusters=k, random_state=42, n_init=10).fit(lab_n)
	km = KMeansRows64(eng, torch.from_numpy(np.ascontiguousarray(lab_n)).to(eng.dev))
	tol = float(np.var(lab_n, axis=0).mean() * 1e-4)
	_, inits = eng.kmeanspp_seeds(None, None, k, 10, 42, rows=km.rows)
	from image_segmenter_b200 import color_simplify as cs
	host_idx = cs._seed_kmeans_plusplus(lab_n, k, 10)
	for j, (a, hi) in enumerate(zip(inits, host_idx)):
		same = np.allclose(a, lab_n[hi])
		labd, ind = km._run(a, 300, tol)
		print("  init", j, "seeds equal sklearn's:", same, "device inertia", ind, "labels used", len(torch.unique(labd)))
	mine = km.fit_predict(k)
	print("  sklearn inertia", ref.inertia_, "labels used", len(np.unique(ref.labels_)), "n_iter", ref.n_iter_,
	      "| device labels used", len(np.unique(mine)), "equal labels:", np.array_equal(mine, ref.labels_),
	      "mismatch", int((mine != ref.labels_).sum()))
