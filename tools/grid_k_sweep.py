"""Exact-label Lloyd iteration at 64 MP (uniform-random sRGB -> LAB planes), labels written: grid-filtered assignment
vs the full walk for K = 16, 32, 64.  python tools/grid_k_sweep.py > gpurun_out/grid_k_sweep.json"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import _ffi
from image_segmenter_b200.engine import get_engine
from image_segmenter_b200.sharded import make_gpu_lloyd

eng = get_engine(0)
n = 8192 * 8192
g = torch.Generator(device=eng.dev)
g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
rgba[:, 3] = 255
planes = eng.rgba_to_lab(rgba)
del rgba
labels = torch.empty(n, dtype=torch.uint8, device=eng.dev)
out = {}
for K in (int(a) for a in (sys.argv[1:] or ["16", "32", "64"])):
	idx = torch.from_numpy(np.random.default_rng(1).choice(n, K, replace=False)).to(eng.dev)
	C0 = np.ascontiguousarray(planes[:, idx].T.double().cpu().numpy())
	res = {}
	for mode, box, pol, exact in (("grid_exact", _ffi.CS_LAB_BOX, 1, True), ("walk_exact", None, 0, True), ("walk_fast", None, 0, False)):
		if mode not in os.environ.get("CS_SWEEP_MODES", "grid_exact,walk_exact,walk_fast").split(","):
			continue
		drv = make_gpu_lloyd(eng, planes, n, K, labels=labels, exact=exact, box=box, grid_policy=pol)
		drv.set_centers(C0)
		for _ in range(12):
			drv.iterate()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		torch.cuda.synchronize()
		e0.record()
		for _ in range(20):
			drv.iterate()
		e1.record()
		torch.cuda.synchronize()
		res[mode] = {"ms_per_iteration": round(e0.elapsed_time(e1) / 20, 4), "centres_checksum": float(drv.c[drv.cur].sum().item())}
	out[f"K{K}"] = res
print(json.dumps(out, indent=1))
