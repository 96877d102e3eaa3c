"""Where the grid-filtered assignment starts to pay: exact-label Lloyd iteration, grid path (policy 1 = wherever
eligible) against the full walk, over shard sizes and K.  python tools/grid_size_sweep.py > gpurun_out/grid_size_sweep.json"""
import json
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
g = torch.Generator(device=eng.dev)
out = {}
for n in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 3840 * 2160, 1 << 24, 1 << 25):
	g.manual_seed(3)
	rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
	rgba[:, 3] = 255
	planes = eng.rgba_to_lab(rgba)
	del rgba
	labels = torch.empty(n, dtype=torch.uint8, device=eng.dev)
	for K in (int(a) for a in (sys.argv[1:] or ["12", "16", "32", "64"])):
		idx = torch.from_numpy(np.random.default_rng(1).choice(n, K, replace=False)).to(eng.dev)
		C0 = np.ascontiguousarray(planes[:, idx].T.double().cpu().numpy())
		res = {}
		for mode, box, pol in (("grid", _ffi.CS_LAB_BOX, 1), ("walk", None, 0)):
			drv = make_gpu_lloyd(eng, planes, n, K, labels=labels, exact=True, box=box, grid_policy=pol)
			drv.set_centers(C0)
			for _ in range(12):
				drv.iterate()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			torch.cuda.synchronize()
			e0.record()
			for _ in range(30):
				drv.iterate()
			e1.record()
			torch.cuda.synchronize()
			res[mode] = round(e0.elapsed_time(e1) / 30 * 1e3, 2)
		out[f"n{n}_K{K}"] = res
	del planes, labels
print(json.dumps(out, indent=1))
