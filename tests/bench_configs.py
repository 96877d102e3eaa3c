"""Measure the five BASELINE.json configs on one B200 next to the reference's CPU path.

Not the driver's bench (that is bench.py = the headline metric); this fills in the per-config table of
DESIGN.md / profiles/.  GPU numbers are wall-clock through the public drop-in API (host arrays in, host
arrays out, so H2D/D2H are inside) unless marked "device".  The CPU side is the oracle pipeline, i.e. the
same scikit-learn / Pillow / OpenCV calls the reference makes, on the same input (or a stated subsample).

usage: python tests/bench_configs.py [--quick] > gpurun_out/configs.json
"""
import argparse
import json
import os
import sys
import time
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
warnings.simplefilter("ignore")

import torch  # noqa: E402

from image_segmenter_b200 import color_simplify as cs  # noqa: E402
from image_segmenter_b200.engine import KMeansGPU, get_engine  # noqa: E402
from oracle import cpu_baseline as cb  # noqa: E402
from oracle import pipeline as op  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
eng = get_engine(0)
out = {"host_cpus": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}


def wall(fn, reps=3):
	best = None
	for _ in range(reps):
		torch.cuda.synchronize()
		t0 = time.perf_counter()
		r = fn()
		torch.cuda.synchronize()
		dt = time.perf_counter() - t0
		best = dt if best is None or dt < best else best
	return best, r


def rand_rgba(seed, h, w):
	rng = np.random.default_rng(seed)
	return np.dstack([rng.integers(0, 256, (h, w, 3), dtype=np.uint8), np.full((h, w), 255, np.uint8)])


def device_lloyd_rate(planes, n, K, exact, iters=20):
	rng = np.random.default_rng(1)
	idx = torch.from_numpy(rng.choice(n, K, replace=False)).to(eng.dev)
	C0 = np.ascontiguousarray(planes[:, idx].T.double().cpu().numpy())
	km = KMeansGPU(eng, "f32", n, planes=planes, exact=exact)
	c = [torch.from_numpy(C0).to(eng.dev), torch.zeros((K, 3), dtype=torch.float64, device=eng.dev)]
	sums, counts = torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev)
	stats = torch.zeros(4, dtype=torch.float64, device=eng.dev)
	for _ in range(5):
		km._step(c[0], K, sums, counts, labels=km.labels, d_cout=c[1], d_stats=stats)
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	torch.cuda.synchronize()
	e0.record()
	cur = 0
	for _ in range(iters):
		km._step(c[cur], K, sums, counts, labels=km.labels, d_cout=c[cur ^ 1], d_stats=stats)
		cur ^= 1
	e1.record()
	torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / iters
	return {"ms_per_iter": round(ms, 4), "mpix_s": round(n / ms / 1e3, 1), "hbm_frac_13B": round(13.0 * n / (ms * 1e-3) / 1e9 / 6549.8, 4)}


# ---- C1: K-means 16 colours on the 1024^2 few-colour working image (stand-in for working_image_cleaned.bmp) ----
rng = np.random.default_rng(13)
pal = np.array([[0, 0, 0], [3, 8, 4], [154, 202, 176], [38, 115, 73], [234, 147, 51], [184, 187, 158], [111, 248, 67],
                [170, 85, 127], [252, 253, 254]], dtype=np.uint8)
idx = rng.choice(9, size=(1024, 1024), p=[0.5, 0.39, 0.03, 0.02, 0.02, 0.01, 0.01, 0.01, 0.01])
img1 = np.dstack([pal[idx], np.full((1024, 1024), 255, np.uint8)])
cs.simplify_colors_kmeans(img1, 16)  # warm
t_gpu, (o1, p1) = wall(lambda: cs.simplify_colors_kmeans(img1, 16))
t0 = time.perf_counter()
ro, rp = op.kmeans_rgb(img1, 16, intended_remap=True)
t_cpu = time.perf_counter() - t0
out["C1_kmeans16_1024sq_fewcolours"] = {
	"gpu_s": round(t_gpu, 4), "cpu_reference_s": round(t_cpu, 4), "speedup": round(t_cpu / t_gpu, 1),
	"palette_rows": int(len(p1)), "palette_equal_or_plus1": bool((((p1.astype(int) - rp.astype(int)) == 0) | ((p1.astype(int) - rp.astype(int)) == 1)).all()) if p1.shape == rp.shape else False,
	"note": "same 9-colour distribution as app/working_image_cleaned.bmp (the BMP itself is not redistributed); K collapses to 7; "
	        "GPU time includes the device k-means++ seeding of 10 inits (host keeps the RandomState stream) and H2D/D2H"}

# ---- C2: perceptual LAB clustering k=16 on 3840x2160 ----
img2 = rand_rgba(2, 2160, 3840)
np.random.seed(2)
cs.simplify_colors_perceptual_fast(img2, 16)
np.random.seed(2)
t_gpu, (o2, p2) = wall(lambda: (np.random.seed(2), cs.simplify_colors_perceptual_fast(img2, 16))[1])
np.random.seed(2)
t0 = time.perf_counter()
ro2, rp2 = op.perceptual_fast(img2, 16)
t_cpu = time.perf_counter() - t0
d2 = eng.upload_rgba(img2)
planes2 = eng.rgba_to_lab(d2)
out["C2_perceptual_fast16_4k"] = {
	"gpu_s": round(t_gpu, 4), "cpu_reference_s": round(t_cpu, 4), "speedup": round(t_cpu / t_gpu, 1),
	"palette_equal": bool(np.array_equal(p2, rp2)), "pixels_differing": int((o2 != ro2).any(axis=2).sum()),
	"device_lloyd_k16_fast": device_lloyd_rate(planes2, d2.shape[0], 16, False),
	"device_lloyd_k16_exact": device_lloyd_rate(planes2, d2.shape[0], 16, True)}
del planes2, d2

# ---- C3: LAB k-means on 64 MP, K=64 and K=16 (1 GPU here; N>1 via bench.py under torchrun) ----
n3 = 8192 * 8192
g = torch.Generator(device=eng.dev)
g.manual_seed(3)
rgba3 = torch.randint(0, 256, (n3, 4), dtype=torch.uint8, device=eng.dev, generator=g)
planes3 = eng.rgba_to_lab(rgba3)
del rgba3
c3 = {}
for K in (8, 16, 32, 64, 128, 256):
	c3[f"k{K}_fast"] = device_lloyd_rate(planes3, n3, K, False, iters=10 if K <= 64 else 3)
for K in (16, 64):
	c3[f"k{K}_exact"] = device_lloyd_rate(planes3, n3, K, True, iters=10)
# natural-image proxy: 6 colour blobs of 512 x 512 tiles + N(0, 12^2) noise, clipped
g.manual_seed(33)
tiles = torch.randint(0, 6, (16, 16), device=eng.dev, generator=g)
cent6 = torch.randint(30, 256, (6, 3), device=eng.dev, generator=g).float()
which3 = tiles.repeat_interleave(512, 0).repeat_interleave(512, 1).reshape(-1)
blob = torch.empty((n3, 4), dtype=torch.uint8, device=eng.dev)
for ch in range(3):
	blob[:, ch] = (cent6[which3, ch] + torch.randn(n3, device=eng.dev, generator=g) * 12.0).clamp_(0, 255).to(torch.uint8)
blob[:, 3] = 255
planes_b = eng.rgba_to_lab(blob)
del blob, which3
c3["blobby_k16_fast"] = device_lloyd_rate(planes_b, n3, 16, False, iters=10)
c3["blobby_k16_exact"] = device_lloyd_rate(planes_b, n3, 16, True, iters=10)
del planes_b
Xs = cb.make_lab_sample(1 << 23, 3)
for K in (16, 64):
	C0 = Xs[np.random.default_rng(0).choice(len(Xs), K, replace=False)]
	secs, kind, threads, _ = cb.time_lloyd_iterations(Xs, C0, 2)
	c3[f"cpu_sklearn_k{K}_mpix_s"] = round(len(Xs) * 2 / secs / 1e6, 1)
	c3["cpu_threads"] = threads
out["C3_lab_kmeans_64mp_device"] = c3
del planes3
torch.cuda.empty_cache()

# ---- C4: batch of 1920x1080 images, k=8 RGB k-means (images are independent: partitioned, no collective) ----
nimg = 16 if args.quick else 128  # one GPU's share of the 1024-image batch
n4 = 1920 * 1080
g.manual_seed(4)
batch = torch.randint(0, 256, (nimg, n4, 4), dtype=torch.uint8, device=eng.dev, generator=g)
batch[:, :, 3] = 255
K = 8
iters = 20
from image_segmenter_b200.batch import kmeans_rgb_batch  # noqa: E402

inits = (batch[:, :K, :3].double() + torch.arange(K, device=eng.dev, dtype=torch.float64)[None, :, None] * 0.01).cpu().numpy()


def run_batch():
	return kmeans_rgb_batch(batch, K, inits, iters - 1, exact=False)  # iters-1 iterations + the final E-step = iters launches


run_batch()
t_gpu, _ = wall(run_batch, reps=3)
sub = batch[:2].cpu().numpy()
t_cpu = 0.0
for i in range(2):
	X = sub[i][:, :3].astype(np.float64)
	t_cpu += cb.time_lloyd_iterations(X, X[:K] + 0.01 * np.arange(K)[:, None], iters)[0] / 2
out["C4_batch_1080p_k8"] = {
	"images_on_this_gpu": nimg, "iterations": iters, "gpu_s": round(t_gpu, 4),
	"gpu_mpix_s_per_iter": round(nimg * n4 * iters / t_gpu / 1e6, 1),
	"cpu_sklearn_s_per_image": round(t_cpu, 4), "cpu_mpix_s_per_iter": round(n4 * iters / t_cpu / 1e6, 1),
	"note": "cs_lloyd_iter_rgba8_batched: ONE launch per iteration for all images of the GPU's share (packed RGBA8 "
	        "features, 5 B/px, labels written in the last pass)"}
del batch
torch.cuda.empty_cache()

# ---- C5: median-cut / octree / posterize, 256 colours, 16 MP, bit-exact ----
img5 = rand_rgba(5, 1024 if args.quick else 4096, 4096)
c5 = {}
for name, fn, ofn in (("median_cut", cs.simplify_colors_median_cut, lambda i, k: op.median_cut(i, k)),
                      ("octree", cs.simplify_colors_octree, lambda i, k: op.median_cut(i, k, power_of_two=False)),
                      ("threshold", cs.simplify_colors_threshold, op.threshold)):
	fn(img5, 256)
	t_gpu, (o5, p5) = wall(lambda: fn(img5, 256), reps=2)
	t0 = time.perf_counter()
	if name == "threshold":
		ro5, rp5 = ofn(img5, 256)
	else:
		from PIL import Image

		im = Image.fromarray(np.ascontiguousarray(img5[:, :, :3])).quantize(colors=256, method=Image.Quantize.MEDIANCUT)
		rp5 = np.array(im.getpalette()).reshape(-1, 3)[:256]
		ro5 = np.dstack([np.array(im.convert("RGB")), img5[:, :, 3]])
	t_cpu = time.perf_counter() - t0
	c5[name] = {"gpu_s": round(t_gpu, 4), "cpu_reference_s": round(t_cpu, 3), "speedup": round(t_cpu / t_gpu, 1),
	            "image_bit_exact": bool(np.array_equal(o5, ro5)), "palette_bit_exact": bool(np.array_equal(p5, rp5)),
	            "mpix_s_gpu": round(img5.shape[0] * img5.shape[1] / t_gpu / 1e6, 1)}
t_gpu, st = wall(lambda: cs.get_color_statistics(img5), reps=2)
t0 = time.perf_counter()
rs = op.statistics(img5)
t_cpu = time.perf_counter() - t0
c5["get_color_statistics"] = {"gpu_s": round(t_gpu, 4), "cpu_reference_s": round(t_cpu, 3), "speedup": round(t_cpu / t_gpu, 1),
                              "unique_equal": st["total_unique_colors"] == rs["total_unique_colors"]}
out["C5_integer_paths_16mp_k256"] = c5

# ---- next row: analyze_regions on a colour-simplified 16 MP image ----
from image_segmenter_b200 import region_cleanup as rcl  # noqa: E402
from oracle import regions as oreg  # noqa: E402

side = 1024 if args.quick else 4096
rng = np.random.default_rng(6)
cent = rng.integers(30, 256, (8, 3))
yy, xx = np.mgrid[0:side, 0:side]
which = ((yy // 257) * 3 + (xx // 301) + ((yy * 7 + xx * 3) // 1999)) % 8
img6 = np.dstack([cent[which].astype(np.uint8), np.full((side, side), 255, np.uint8)])
spk = rng.random((side, side)) < 0.002  # isolated speckles: the small regions the clean-up step hunts
img6[spk, :3] = cent[rng.integers(0, 8, int(spk.sum()))]
rcl.analyze_regions(img6)
t_gpu, r6 = wall(lambda: rcl.analyze_regions(img6), reps=2)
t0 = time.perf_counter()
o6 = oreg.analyze_regions(img6)
t_cpu = time.perf_counter() - t0
out["N4_analyze_regions_16mp_8colours"] = {
	"gpu_s": round(t_gpu, 4), "cpu_reference_s": round(t_cpu, 3), "speedup": round(t_cpu / t_gpu, 1),
	"total_regions": r6["total_regions"], "equal_to_reference": bool(r6["region_sizes"] == o6["region_sizes"] and
	[a["label"] for a in r6["all_regions"]] == [a["label"] for a in o6["all_regions"]]),
	"note": "GPU time includes the download of the 8 per-colour label / mask arrays the reference API returns (5 B/px per colour)"}
print(json.dumps(out, indent=1))
