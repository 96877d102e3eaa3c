"""Exact-label Lloyd iteration at 64 MP, K=16: grid-filtered assignment vs the full walk, on the uniform-random
benchmark image and on a spatially coherent one (6 colour blobs in 512 x 512 tiles + N(0, 12^2) noise).
python tools/grid_vs_walk.py > gpurun_out/grid_vs_walk.json"""
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
n = 8192 * 8192
g = torch.Generator(device=eng.dev)
out = {}


def images():
	g.manual_seed(3)
	rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
	rgba[:, 3] = 255
	yield "uniform_random", rgba
	del rgba
	g.manual_seed(33)
	tiles = torch.randint(0, 6, (16, 16), device=eng.dev, generator=g)
	cent6 = torch.randint(30, 256, (6, 3), device=eng.dev, generator=g).float()
	which = tiles.repeat_interleave(512, 0).repeat_interleave(512, 1).reshape(-1)
	blob = torch.empty((n, 4), dtype=torch.uint8, device=eng.dev)
	for ch in range(3):
		blob[:, ch] = (cent6[which, ch] + torch.randn(n, device=eng.dev, generator=g) * 12.0).clamp_(0, 255).to(torch.uint8)
	blob[:, 3] = 255
	yield "blobs_plus_noise", blob


for name, rgba in images():
	planes = eng.rgba_to_lab(rgba)
	idx = torch.from_numpy(np.random.default_rng(1).choice(n, 16, replace=False)).to(eng.dev)
	C0 = np.ascontiguousarray(planes[:, idx].T.double().cpu().numpy())
	labels = torch.empty(n, dtype=torch.uint8, device=eng.dev)
	res = {}
	for mode, box, exact in (("grid_exact", _ffi.CS_LAB_BOX, True), ("walk_exact", None, True), ("walk_fast", None, False)):
		drv = make_gpu_lloyd(eng, planes, n, 16, labels=labels, exact=exact, box=box)
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
	out[name] = res
	del planes
print(json.dumps(out, indent=1))
