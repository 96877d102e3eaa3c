"""Fixed cost per Lloyd launch: time batches queued through cs_lloyd_run_f32 (no host work between
launches) for shrinking images (development probe)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200.engine import get_engine

eng = get_engine(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
SIZES = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (4096, 65536, 262144, 1 << 20, 1920 * 1080, 3840 * 2160, 4096 * 4096, 8192 * 8192)
for n in SIZES:
	g = torch.Generator(device=eng.dev); g.manual_seed(1)
	rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
	planes = eng.rgba_to_lab(rgba)
	a = planes[:, :K].T.double().contiguous() + 0.01
	b = torch.zeros_like(a)
	sums, counts, stats = (torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev),
	                       torch.zeros(4, dtype=torch.float64, device=eng.dev))
	ctl = torch.tensor([0.0, 0.0, -1.0, 0.0], dtype=torch.float64, device=eng.dev)
	run = lambda m: eng._call("cs_lloyd_run_f32", planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), n, a.data_ptr(),
	                          b.data_ptr(), K, sums.data_ptr(), counts.data_ptr(), stats.data_ptr(), 31400.0, 0, m, ctl.data_ptr())
	run(10)
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record(); run(100); e1.record(); torch.cuda.synchronize()
	us = e0.elapsed_time(e1) * 10
	ideal = 13.0 * n / 6549.8e9 * 1e6 / 0.566
	ts = np.zeros(64, dtype=np.uint64)
	eng.ctx.lib.cs_debug_scratch(eng.ctx.handle, ts.ctypes.data)
	st = ts[16:23].astype(np.int64)
	if st[0]:
		print("   phases ns: prologue %d | main %d | fold+partial %d | fence+atomic %d | combine %d | finalize %d | total %d" % (
			st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[6] - st[0]))
	print(f"n={n:9d} K={K}: {us:7.2f} us/iter  (pure streaming part at 56.6% HBM would be {ideal:6.2f} us)  halt={ctl[0].item()} iters={ctl[1].item()}", flush=True)
