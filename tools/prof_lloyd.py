"""Run the Lloyd step a few times on synthetic LAB planes (for ncu / clock sampling).
usage: prof_lloyd.py N K VARIANT FLAGS ITERS"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
lib = C.CDLL(str(ROOT / "image_segmenter_b200/_lib/libcolorsimplify.so"))
lib.cs_last_error.restype = C.c_char_p
vp = C.c_void_p
n, K, variant, fl, iters = (int(a) for a in sys.argv[1:6])
ctx = vp()
assert lib.cs_ctx_create(0, C.byref(ctx)) == 0
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
planes = [torch.rand(n, device=dev, generator=g) * s + o for s, o in ((100, 0), (185, -90), (200, -105))]
idx = torch.randint(0, n, (K,), device=dev, generator=g)
Cn = torch.stack([p[idx] for p in planes], 1).double().contiguous()
d_lab = torch.empty(n, dtype=torch.uint8, device=dev)
d_sums = torch.zeros(K * 3, dtype=torch.float64, device=dev); d_cnt = torch.zeros(K, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
flags = fl | (variant << 8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters + 3):
	if i == 3:
		e0.record()
	rc = lib.cs_lloyd_step_f32(ctx, vp(planes[0].data_ptr()), vp(planes[1].data_ptr()), vp(planes[2].data_ptr()),
	                           C.c_int64(n), vp(Cn.data_ptr()), K, vp(d_lab.data_ptr()), vp(d_sums.data_ptr()),
	                           vp(d_cnt.data_ptr()), None, C.c_double(31400.0), flags, vp(st))
	assert rc == 0, lib.cs_last_error()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"n={n} K={K} variant={variant} flags={fl} ms={ms:.4f} GB/s={13.0 * n / ms / 1e6:.1f} count={d_cnt.sum().item():.0f}")
