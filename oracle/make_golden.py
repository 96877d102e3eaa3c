"""Generate tests/golden/*.npz from the UNMODIFIED reference (authoring container only).

TEST INFRASTRUCTURE.  Imports /root/reference/app/processing/color_simplify.py as is
(`sys.path.insert(0, "/root/reference/app")`), runs its entry points on small seeded synthetic
RGBA images and stores inputs + outputs as compressed fixtures.  scikit-image is not installed,
so the four entry points that import it run with `oracle.lab` injected as `skimage.color`
(oracle/lab.py: install_skimage_stub) — those fixtures pin the reference's control flow and
sklearn fits, not skimage's arithmetic.  Also stores direct outputs of sklearn's
`lloyd_iter_chunked_dense` / `KMeans` (the kernel the reference's KMeans.fit runs) on seeded LAB
data.  /root/reference does not exist on the GPU box: tests read only the committed .npz files.

usage: PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "golden"
REF_APP = "/root/reference/app"


def synth_images():
	"""Small seeded RGBA inputs (kept tiny so the fixtures stay small)."""
	rng = np.random.default_rng(11)
	h, w = 80, 96
	cent = rng.integers(30, 256, (6, 3))
	which = (np.add.outer(np.arange(h) // 14, np.arange(w) // 17) + rng.integers(0, 2, (h, w))) % 6
	rgb = np.clip(cent[which] + rng.normal(0, 9, (h, w, 3)), 0, 255).astype(np.uint8)
	alpha = np.full((h, w), 255, np.uint8)
	alpha[:10, :] = 0
	alpha[10:14, :] = 100
	alpha[14:18, 5:40] = 200
	rgb[60:, 70:] = rng.integers(0, 12, (20, 26, 3))  # a dark corner for the brightness filters
	blobby = np.dstack([rgb, alpha])

	rng = np.random.default_rng(12)
	uni = np.dstack([rng.integers(0, 256, (64, 64, 3), dtype=np.uint8), np.full((64, 64), 255, np.uint8)])

	rng = np.random.default_rng(13)  # few colours, mostly dark: the character of working_image_cleaned.bmp
	pal = np.array([[0, 0, 0], [3, 8, 4], [154, 202, 176], [38, 115, 73], [234, 147, 51], [184, 187, 158],
	                [111, 248, 67], [170, 85, 127], [252, 253, 254]], dtype=np.uint8)
	idx = rng.choice(9, size=(96, 96), p=[0.5, 0.39, 0.03, 0.02, 0.02, 0.01, 0.01, 0.01, 0.01])
	few = np.dstack([pal[idx], np.full((96, 96), 255, np.uint8)])
	return {"blobby": blobby, "uniform": uni, "fewcolors": few}


def adaptive_distance_golden():
	"""tests/golden/reference_adaptive_distance.npz: simplify_colors_adaptive_distance of the unmodified
	reference where it runs, and the IndexError it raises where its merge branch is hit."""
	sys.dont_write_bytecode = True
	sys.path.insert(0, str(ROOT))
	from oracle import lab as olab

	olab.install_skimage_stub()
	sys.path.insert(0, REF_APP)
	from processing import color_simplify as ref

	imgs = synth_images()
	store = {}
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		for name, k in (("blobby", 3), ("blobby", 6), ("blobby", 8), ("fewcolors", 8), ("uniform", 6), ("fewcolors", 3)):
			try:
				out, pal = ref.simplify_colors_adaptive_distance(imgs[name], k)
				store[f"{name}__ad_{k}__rgba"], store[f"{name}__ad_{k}__palette"] = out, pal
			except IndexError as e:
				store[f"{name}__ad_{k}__indexerror"] = np.array([1])
	np.savez_compressed(OUT / "reference_adaptive_distance.npz", **store)
	print("wrote reference_adaptive_distance.npz", sorted(store))


def regions_golden():
	"""tests/golden/reference_regions.npz: analyze_regions of the unmodified reference
	(app/processing/region_cleanup.py) on colour-simplified versions of the seeded images."""
	sys.dont_write_bytecode = True
	sys.path.insert(0, REF_APP)
	from processing import region_cleanup as ref

	g = np.load(OUT / "reference_entry_points.npz")
	cases = {"blobby_t8": g["blobby__threshold_8__rgba"], "blobby_mc8": g["blobby__median_cut_8__rgba"],
	         "fewcolors": g["in_fewcolors"], "uniform_t2": g["uniform__threshold_2__rgba"]}
	store = {}
	for name, img in cases.items():
		store[f"in_{name}"] = img
		for conn in (8, 4):
			r = ref.analyze_regions(img, 100, conn)
			tag = f"{name}__c{conn}"
			store[f"{tag}__summary"] = np.array([r["total_regions"], r["small_regions"], r["largest_region_size"],
			                                      r["smallest_region_size"]], dtype=np.int64)
			store[f"{tag}__colors"] = np.array(r["region_colors"], dtype=np.uint8).reshape(-1, 3)
			store[f"{tag}__sizes"] = np.array(r["region_sizes"], dtype=np.int64)
			store[f"{tag}__labels"] = np.array([a["label"] for a in r["all_regions"]], dtype=np.int64)
			store[f"{tag}__bbox"] = np.array([a["bbox"] for a in r["all_regions"]], dtype=np.int64).reshape(-1, 4)
			ucol = np.unique(store[f"{tag}__colors"], axis=0)
			lab_stack, mask_stack = [], []
			for c in ucol:
				a = next(a for a in r["all_regions"] if tuple(a["color"]) == tuple(c))
				lab_stack.append(a["labels"])
				mask_stack.append(a["color_mask"])
			store[f"{tag}__label_images"] = np.array(lab_stack, dtype=np.int32)
			store[f"{tag}__mask_images"] = np.array(mask_stack, dtype=np.uint8)
			store[f"{tag}__dist"] = np.array([r["size_distribution"].get(k, 0) for k in ("< 50", "50-99", "100-199", "200-499", "500+")])
	np.savez_compressed(OUT / "reference_regions.npz", **store)
	print("wrote reference_regions.npz", (OUT / "reference_regions.npz").stat().st_size)


def edge_images():
	"""Degenerate and ragged inputs: the cases the reference guards with early returns or fallbacks."""
	rng = np.random.default_rng(21)
	op = lambda a: np.dstack([a, np.full(a.shape[:2], 255, np.uint8)])
	tiny = op(rng.integers(0, 256, (3, 5, 3), dtype=np.uint8))
	onepx = op(np.array([[[200, 100, 50]]], dtype=np.uint8))
	ragged = np.dstack([rng.integers(0, 256, (7, 13, 3), dtype=np.uint8),
	                    rng.choice(np.array([0, 1, 128, 129, 255], dtype=np.uint8), size=(7, 13))])
	transparent = np.dstack([rng.integers(0, 256, (6, 5, 3), dtype=np.uint8), np.zeros((6, 5), np.uint8)])
	dark = op(rng.integers(0, 3, (8, 9, 3), dtype=np.uint8))            # brightness < 10 everywhere
	mid = op(rng.integers(12, 28, (8, 9, 3), dtype=np.uint8))           # 10 < brightness < 30: the fallback threshold
	two = op(np.where(rng.random((10, 11, 1)) < 0.4, np.array([[[220, 40, 40]]]), np.array([[[40, 60, 200]]])).astype(np.uint8))
	return {"tiny": tiny, "onepx": onepx, "ragged": ragged, "transparent": transparent, "dark": dark, "midbright": mid,
	        "twocolors": two}


def edge_cases_golden():
	"""tests/golden/reference_edge_cases.npz: every entry point of the unmodified reference on edge_images();
	an exception is stored as its type name."""
	sys.dont_write_bytecode = True
	sys.path.insert(0, str(ROOT))
	from oracle import lab as olab

	olab.install_skimage_stub()
	sys.path.insert(0, REF_APP)
	from processing import color_simplify as ref

	cp = np.array([[250, 10, 10], [10, 240, 30], [20, 30, 230], [128, 128, 128]], dtype=np.uint8)
	calls = {
		"kmeans_8": lambda im: ref.simplify_colors_kmeans(im, 8),
		"median_cut_8": lambda im: ref.simplify_colors_median_cut(im, 8),
		"octree_5": lambda im: ref.simplify_colors_octree(im, 5),
		"threshold_8": lambda im: ref.simplify_colors_threshold(im, 8),
		"threshold_8_noalpha": lambda im: ref.simplify_colors_threshold(im, 8, preserve_alpha=False),
		"hsv_4": lambda im: ref.simplify_colors_hsv_clustering(im, 4),
		"custom_rgb": lambda im: ref.simplify_colors_custom_palette(im, cp, True, "rgb"),
		"custom_lab": lambda im: ref.simplify_colors_custom_palette(im, cp, True, "lab"),
		"custom_hsv_noalpha": lambda im: ref.simplify_colors_custom_palette(im, cp, False, "hsv"),
		"perceptual_fast_4": lambda im: ref.simplify_colors_perceptual_fast(im, 4),
		"perceptual_3": lambda im: ref.simplify_colors_perceptual(im, 3, max_samples=2000),
	}
	store = {"custom_palette_in": cp}
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		for name, img in edge_images().items():
			store[f"in_{name}"] = img
			for tag, fn in calls.items():
				np.random.seed(7)
				try:
					out, pal = fn(img)
					store[f"{name}__{tag}__rgba"] = np.asarray(out)
					store[f"{name}__{tag}__palette"] = np.asarray(pal)
					store[f"{name}__{tag}__same_object"] = np.array([out is img])
				except Exception as e:  # noqa: BLE001 - the type is the fixture
					store[f"{name}__{tag}__raises"] = np.array([type(e).__name__])
			st = ref.get_color_statistics(img)
			store[f"{name}__stats"] = np.array([st["total_unique_colors"], st["non_transparent_pixels"], *st["rgb_mean"],
			                                   *st["rgb_std"]], dtype=np.float64)
	np.savez_compressed(OUT / "reference_edge_cases.npz", **store)
	raised = sorted(k for k in store if k.endswith("__raises"))
	print("wrote reference_edge_cases.npz", (OUT / "reference_edge_cases.npz").stat().st_size, "bytes;", len(raised), "calls raise:",
	      [(k, str(store[k][0])) for k in raised])


def main():
	sys.dont_write_bytecode = True
	sys.path.insert(0, str(ROOT))
	from oracle import lab as olab

	olab.install_skimage_stub()
	sys.path.insert(0, REF_APP)
	from processing import color_simplify as ref  # the unmodified reference module

	OUT.mkdir(parents=True, exist_ok=True)
	imgs = synth_images()
	store = {}
	for name, img in imgs.items():
		store[f"in_{name}"] = img

	def put(tag, res):
		out, pal = res
		store[f"{tag}__rgba"] = np.asarray(out)
		store[f"{tag}__palette"] = np.asarray(pal)

	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		for name, img in imgs.items():
			for k in (5, 16):
				put(f"{name}__kmeans_{k}", ref.simplify_colors_kmeans(img, k))
			for k in (8, 16, 100):
				put(f"{name}__median_cut_{k}", ref.simplify_colors_median_cut(img, k))
			for k in (6, 16, 256):
				put(f"{name}__octree_{k}", ref.simplify_colors_octree(img, k))
			for k in (2, 8, 16, 256):
				put(f"{name}__threshold_{k}", ref.simplify_colors_threshold(img, k))
			put(f"{name}__threshold_8_noalpha", ref.simplify_colors_threshold(img, 8, preserve_alpha=False))
			put(f"{name}__hsv_6", ref.simplify_colors_hsv_clustering(img, 6))
			cp = np.array([[250, 10, 10], [10, 240, 30], [20, 30, 230], [128, 128, 128], [0, 0, 0], [255, 255, 255],
			               [128, 128, 128]], dtype=np.uint8)
			for metric in ("rgb", "hsv", "lab"):
				put(f"{name}__custom_{metric}", ref.simplify_colors_custom_palette(img, cp, True, metric))
			put(f"{name}__custom_lab_noalpha", ref.simplify_colors_custom_palette(img, cp, False, "lab"))
			np.random.seed(7)
			put(f"{name}__perceptual_fast_6", ref.simplify_colors_perceptual_fast(img, 6))
			np.random.seed(7)
			put(f"{name}__perceptual_5", ref.simplify_colors_perceptual(img, 5, max_samples=2000))
			st = ref.get_color_statistics(img)
			store[f"{name}__stats"] = np.array([st["total_unique_colors"], st["non_transparent_pixels"], *st["rgb_mean"],
			                                   *st["rgb_std"]], dtype=np.float64)
			put(f"{name}__adaptive_kmeans_4", ref.simplify_colors_adaptive(img, 4, True, "kmeans"))
		store["custom_palette_in"] = cp

		# the real fixture of BASELINE config 1: known answers of the unmodified reference, and the input itself as
		# a 9-entry colour table + a 1024 x 1024 index map (tests/golden/working_image_cleaned.npz, ~60 KB compressed)
		bmp = Path(REF_APP) / "working_image_cleaned.bmp"
		if bmp.exists():
			from PIL import Image

			im = np.array(Image.open(bmp).convert("RGB"))
			cols, inv = np.unique(im.reshape(-1, 3), axis=0, return_inverse=True)
			assert len(cols) <= 256
			np.savez_compressed(OUT / "working_image_cleaned.npz", colours=cols.astype(np.uint8),
			                    index=inv.reshape(im.shape[:2]).astype(np.uint8))
			rgba = np.dstack([im, np.full(im.shape[:2], 255, np.uint8)])
			store["bmp__kmeans_16__palette"] = ref.simplify_colors_kmeans(rgba, 16)[1]
			store["bmp__median_cut_16__palette"] = ref.simplify_colors_median_cut(rgba, 16)[1]
			store["bmp__threshold_16__palette"] = ref.simplify_colors_threshold(rgba, 16)[1]
			store["bmp__hsv_16__palette"] = ref.simplify_colors_hsv_clustering(rgba, 16)[1]
			st = ref.get_color_statistics(rgba)
			store["bmp__stats"] = np.array([st["total_unique_colors"], st["non_transparent_pixels"], *st["rgb_mean"],
			                               *st["rgb_std"]], dtype=np.float64)

	np.savez_compressed(OUT / "reference_entry_points.npz", **store)

	# ---- sklearn's own Lloyd kernel on seeded LAB data (fp32-rounded, as the GPU stores it) ----
	from sklearn.cluster import KMeans
	from sklearn.cluster._k_means_lloyd import lloyd_iter_chunked_dense
	from sklearn.utils._openmp_helpers import _openmp_effective_n_threads

	rng = np.random.default_rng(21)
	n, K = 20000, 16
	rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
	lab32 = olab.rgb2lab(rgb.reshape(-1, 1, 3)).reshape(-1, 3).astype(np.float32)
	X = lab32.astype(np.float64)
	C0 = X[rng.choice(n, K, replace=False)].copy()
	cnew = np.zeros_like(C0)
	wts = np.zeros(K)
	labels = np.full(n, -1, np.int32)
	shift = np.zeros(K)
	lloyd_iter_chunked_dense(X, np.ones(n), C0, cnew, wts, labels, shift, _openmp_effective_n_threads())
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		km = KMeans(n_clusters=K, init=C0, n_init=1, max_iter=300, tol=1e-4).fit(X)
	np.savez_compressed(OUT / "sklearn_lloyd.npz", rgb=rgb, lab32=lab32, C0=C0, step_centers=cnew, step_weights=wts,
	                    step_labels=labels, step_shift=shift, fit_centers=km.cluster_centers_,
	                    fit_labels=km.labels_.astype(np.int32), fit_inertia=np.array(km.inertia_),
	                    fit_n_iter=np.array(km.n_iter_))
	print("wrote", sorted(p.name for p in OUT.glob("*.npz")), {p.name: p.stat().st_size for p in OUT.glob("*.npz")})


if __name__ == "__main__":
	if len(sys.argv) > 1 and sys.argv[1] == "adaptive":
		adaptive_distance_golden()
	elif len(sys.argv) > 1 and sys.argv[1] == "regions":
		regions_golden()
	elif len(sys.argv) > 1 and sys.argv[1] == "edge":
		edge_cases_golden()
	else:
		main()
