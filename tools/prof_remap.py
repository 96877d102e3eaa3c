"""Minimal driver for profiling K4 (cs_assign_remap_rgba8) on a 64 MP uniform-random image:
python tools/prof_remap.py [lab|rgb] [K] [policy]   (policy: cs_remap_set_policy, default 0; run plain first, then under ncu -k regex:remap)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import _colorspace as csp
from image_segmenter_b200 import _ffi
from image_segmenter_b200.engine import get_engine

space = sys.argv[1] if len(sys.argv) > 1 else "lab"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = get_engine(0)
policy = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng.set_remap_policy(policy)
n = 8192 * 8192
g = torch.Generator(device=eng.dev)
g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
rgba[:, 3] = 255
pal = np.random.default_rng(0).integers(0, 256, (K, 3), dtype=np.uint8)
feats = csp.rgb2lab_small(pal) if space == "lab" else pal.astype(np.float64)
sp = _ffi.CS_SPACE_LAB if space == "lab" else _ffi.CS_SPACE_RGB
d_c = torch.from_numpy(np.ascontiguousarray(feats)).to(eng.dev)
d_pal = torch.from_numpy(pal).to(eng.dev)
dst = torch.empty_like(rgba)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
	if i == 1:
		e0.record()
	eng._call("cs_assign_remap_rgba8", rgba.data_ptr(), n, sp, eng.lut256.data_ptr(), d_c.data_ptr(), d_pal.data_ptr(), K, 1,
	          dst.data_ptr(), None)
e1.record()
torch.cuda.synchronize()
print(f"K4 {space} K={K} policy={policy}: {e0.elapsed_time(e1) / 3:.4f} ms per 64 MP call")
