"""Device-resident timing of every streaming kernel of the path against the HBM roofline
(algorithmic bytes per pixel of SURVEY.md §8d; CUDA events; inputs larger than L2).
usage: python tools/bench_kernels.py > gpurun_out/kernels.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import _ffi
from image_segmenter_b200.engine import get_engine

eng = get_engine(0)
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
n = 8192 * 8192
g = torch.Generator(device=eng.dev)
g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
rgba[:, 3] = 255
out = {"n_px": n, "peak_gbs": PEAK}


def timeit(name, fn, bytes_per_px, reps=5):
	for _ in range(2):
		fn()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	torch.cuda.synchronize()
	e0.record()
	for _ in range(reps):
		fn()
	e1.record()
	torch.cuda.synchronize()
	ms = e0.elapsed_time(e1) / reps
	gbs = bytes_per_px * n / ms / 1e6
	out[name] = {"ms": round(ms, 4), "mpix_s": round(n / ms / 1e3, 1), "bytes_per_px": bytes_per_px, "gb_s": round(gbs, 1),
	             "hbm_frac": round(gbs / PEAK, 4)}
	print(name, out[name], file=sys.stderr, flush=True)


planes = torch.empty((3, n), dtype=torch.float32, device=eng.dev)
timeit("K1_rgba8_to_lab", lambda: eng._call("cs_rgba8_to_lab", rgba.data_ptr(), n, eng.lut256.data_ptr(), planes[0].data_ptr(),
                                          planes[1].data_ptr(), planes[2].data_ptr()), 16)
hsva = torch.empty_like(rgba)
timeit("K9_rgba8_to_hsv8", lambda: eng._call("cs_rgba8_to_hsv8", rgba.data_ptr(), n, hsva.data_ptr()), 8)
rng = np.random.default_rng(0)
for K in (16, 64):
	pal = rng.integers(0, 256, (K, 3), dtype=np.uint8)
	from image_segmenter_b200 import _colorspace as csp

	d_pal = torch.from_numpy(pal).to(eng.dev)
	dst = torch.empty_like(rgba)
	for space, name, feats in ((_ffi.CS_SPACE_LAB, "lab", csp.rgb2lab_small(pal)), (_ffi.CS_SPACE_RGB, "rgb", pal.astype(np.float64)),
	                           (_ffi.CS_SPACE_HSV, "hsv", csp.rgb2hsv_u8_small(pal).astype(np.float64))):
		d_c = torch.from_numpy(np.ascontiguousarray(feats)).to(eng.dev)
		timeit(f"K4_assign_remap_{name}_k{K}", lambda: eng._call("cs_assign_remap_rgba8", rgba.data_ptr(), n, space, eng.lut256.data_ptr(),
		                                                           d_c.data_ptr(), d_pal.data_ptr(), K, 1, dst.data_ptr(), None), 8, reps=3)
lab = torch.randint(0, 16, (n,), dtype=torch.uint8, device=eng.dev, generator=g)
pal16 = torch.from_numpy(rng.integers(0, 256, (16, 3), dtype=np.uint8)).to(eng.dev)
dst = torch.empty_like(rgba)
timeit("remap_labels", lambda: eng._call("cs_remap_labels_rgba8", rgba.data_ptr(), lab.data_ptr(), n, None, 0, -1, pal16.data_ptr(), 16, 1,
                                        dst.data_ptr()), 9)
bm = eng.bitmap24()
timeit("K7_posterize_step36", lambda: eng._call("cs_posterize_rgba8", rgba.data_ptr(), n, 36, 1, dst.data_ptr(), bm.data_ptr()), 8)
hist = torch.zeros(1 << 24, dtype=torch.int32, device=eng.dev)
timeit("K5_hist_rgb24_uniform", lambda: eng._call("cs_hist_rgb24", rgba.data_ptr(), n, hist.data_ptr()), 4)
low = rgba.clone()
low[:, :3] = (low[:, :3] >> 6) << 6  # 64 distinct colours: the natural-image-like low-entropy case
timeit("K5_hist_rgb24_64colours", lambda: eng._call("cs_hist_rgb24", low.data_ptr(), n, hist.data_ptr()), 4)
acc = torch.zeros(8, dtype=torch.int64, device=eng.dev)
timeit("K8_stats_moments_only", lambda: eng._call("cs_stats_rgba8", rgba.data_ptr(), n, None, acc.data_ptr()), 4)
bm32 = eng.bitmap32()
timeit("K8_stats_with_unique_bitmap", lambda: eng._call("cs_stats_rgba8", rgba.data_ptr(), n, bm32.data_ptr(), acc.data_ptr()), 4)
acc4 = torch.zeros(4, dtype=torch.int64, device=eng.dev)
timeit("K8_mask_stats", lambda: eng._call("cs_mask_stats_rgba8", rgba.data_ptr(), n, 90, None, acc4.data_ptr()), 4)
n16 = 4096 * 4096
sub = rgba[:n16]
o, palm, idx = eng.median_cut(sub, 256, True)
lutq = None
print(json.dumps(out, indent=1))
