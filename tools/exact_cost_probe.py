"""Why does EXACT_TIES cost vary? time fast vs exact for centre sets at different stages of a run and
estimate the near-tie rate on a sample (development probe)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200.engine import KMeansGPU, get_engine

eng = get_engine(0)
n, K = 8192 * 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device=eng.dev); g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
planes = eng.rgba_to_lab(rgba); del rgba
rng = np.random.default_rng(1)
idx = torch.from_numpy(rng.choice(n, K, replace=False)).to(eng.dev)
C = planes[:, idx].T.double().contiguous()
km = {e: KMeansGPU(eng, "f32", n, planes=planes, exact=e) for e in (False, True)}
sums, counts = torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev)
stats = torch.zeros(4, dtype=torch.float64, device=eng.dev)
Cn = torch.zeros_like(C)
sample = planes[:, ::257].T.double()


def timeit(e, c, reps=5):
	for _ in range(2):
		km[e]._step(c, K, sums, counts, labels=km[e].labels, d_cout=Cn, d_stats=stats)
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(reps):
		km[e]._step(c, K, sums, counts, labels=km[e].labels, d_cout=Cn, d_stats=stats)
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) / reps


it = 0
for stage in (0, 1, 2, 5, 10, 20, 50, 100, 200):
	while it < stage:
		km[False]._step(C, K, sums, counts, labels=None, d_cout=Cn, d_stats=stats)
		C, Cn = Cn.clone(), C
		it += 1
	d = ((sample[:, None, :] - C[None]) ** 2).sum(-1)
	top2 = torch.topk(d, 2, dim=1, largest=False).values
	gap = (top2[:, 1] - top2[:, 0])
	cn = float((C ** 2).sum(1).max())
	tau = 1.5 * 2 * 7 * 2.0 ** -24 * ((np.sqrt(31400.0) + np.sqrt(cn)) ** 2 + 1)
	cd = torch.cdist(C, C) + torch.eye(K, device=eng.dev, dtype=torch.float64) * 1e9
	print(f"iter {stage:4d}: fast {timeit(False, C):.4f} ms  exact {timeit(True, C):.4f} ms  tie-rate(gap<=tau={tau:.3f}) "
	      f"{float((gap <= tau).double().mean()):.2e}  min centre dist {float(cd.min()):.2f}  empty {int((counts == 0).sum())}", flush=True)
