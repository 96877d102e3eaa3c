"""Per-iteration time distribution of the fused Lloyd iteration on one GPU, with and without the label store
(development probe for the multi-GPU analysis: every iteration of a sharded run ends at the slowest rank, so
the job pays the sum over iterations of the per-iteration maximum over ranks)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import _ffi
from image_segmenter_b200.engine import get_engine

eng = get_engine(0)
K, n = 16, 8192 * 8192
g = torch.Generator(device=eng.dev); g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
planes = eng.rgba_to_lab(rgba)
del rgba
c = [planes[:, :K].T.double().contiguous() + 0.01, torch.zeros((K, 3), dtype=torch.float64, device=eng.dev)]
sums, counts, stats = (torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev),
                       torch.zeros(4, dtype=torch.float64, device=eng.dev))
labels = torch.empty(n, dtype=torch.uint8, device=eng.dev)
for lab in (None, labels):
	def step(i, chained):
		eng._call("cs_lloyd_iter_f32", planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr(), n, c[i & 1].data_ptr(), K,
		          lab.data_ptr() if lab is not None else None, sums.data_ptr(), counts.data_ptr(), c[(i & 1) ^ 1].data_ptr(),
		          stats.data_ptr(), _ffi.CS_LAB_NORM2_MAX, chained)
	for i in range(50):
		step(i, 0)
	torch.cuda.synchronize()
	ev = [torch.cuda.Event(enable_timing=True) for _ in range(401)]
	ev[0].record()
	for i in range(400):
		step(i, 0)
		ev[i + 1].record()
	torch.cuda.synchronize()
	t = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(400)]) * 1e3
	print(f"labels={'yes' if lab is not None else 'no '}: mean {t.mean():7.2f} us  std {t.std():5.2f}  min {t.min():7.2f}  p50 {np.percentile(t, 50):7.2f}  "
	      f"p90 {np.percentile(t, 90):7.2f}  p99 {np.percentile(t, 99):7.2f}  max {t.max():7.2f}   E[max of 8 draws] {np.mean([t[np.random.default_rng(s).integers(0, 400, 8)].max() for s in range(2000)]):7.2f}", flush=True)
