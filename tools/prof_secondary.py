"""One launch of every secondary kernel on B200-sized inputs, for `ncu --set full -k regex:...`:
K1 rgba8_to_lab, K4 remap_grid (LAB, K=16), K5 hist_rgb24 (uniform colours / 64 colours), palette_map (256
colours), K8 stats with the 2^32-bit bitmap, K7 posterize, ccl_merge (8-colour 16 MP image).
python tools/prof_secondary.py   (plain run first; then under ncu)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from image_segmenter_b200 import _colorspace as csp
from image_segmenter_b200 import _ffi
from image_segmenter_b200.engine import get_engine

eng = get_engine(0)
n = 8192 * 8192
g = torch.Generator(device=eng.dev)
g.manual_seed(3)
rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=eng.dev, generator=g)
rgba[:, 3] = 255
planes = torch.empty((3, n), dtype=torch.float32, device=eng.dev)
dst = torch.empty_like(rgba)
eng._call("cs_rgba8_to_lab", rgba.data_ptr(), n, eng.lut256.data_ptr(), planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr())
pal = np.random.default_rng(0).integers(0, 256, (16, 3), dtype=np.uint8)
d_c = torch.from_numpy(np.ascontiguousarray(csp.rgb2lab_small(pal))).to(eng.dev)
d_pal = torch.from_numpy(pal).to(eng.dev)
eng._call("cs_assign_remap_rgba8", rgba.data_ptr(), n, _ffi.CS_SPACE_LAB, eng.lut256.data_ptr(), d_c.data_ptr(), d_pal.data_ptr(), 16, 1,
          dst.data_ptr(), None)
hist = torch.zeros(1 << 24, dtype=torch.int32, device=eng.dev)
eng._call("cs_hist_rgb24", rgba.data_ptr(), n, hist.data_ptr())
low = rgba.clone()
low[:, :3] = (low[:, :3] >> 6) << 6
hist.zero_()
eng._call("cs_hist_rgb24", low.data_ptr(), n, hist.data_ptr())
del low
acc = torch.zeros(8, dtype=torch.int64, device=eng.dev)
eng._call("cs_stats_rgba8", rgba.data_ptr(), n, eng.bitmap32().data_ptr(), acc.data_ptr())
eng._call("cs_posterize_rgba8", rgba.data_ptr(), n, 36, 1, dst.data_ptr(), eng.bitmap24().data_ptr())
n16 = 4096 * 4096
eng.median_cut(rgba[:n16], 256, True)  # hist + fold + compact + box_sums + palette_map
del planes, dst
torch.cuda.synchronize()
from image_segmenter_b200 import region_cleanup as rcl

side = 4096
rng = np.random.default_rng(6)
cent = rng.integers(30, 256, (8, 3))
yy, xx = np.mgrid[0:side, 0:side]
which = ((yy // 257) * 3 + (xx // 301) + ((yy * 7 + xx * 3) // 1999)) % 8
img6 = np.dstack([cent[which].astype(np.uint8), np.full((side, side), 255, np.uint8)])
rcl.analyze_regions(img6)
torch.cuda.synchronize()
print("ok")
