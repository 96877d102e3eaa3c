"""Variant sweep with a cross-variant checksum (development tool; needs a CS_TUNING_VARIANTS build)."""
import ctypes as C, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
import os
lib = C.CDLL(os.environ.get("COLORSIMPLIFY_LIB", str(ROOT / "image_segmenter_b200/_lib/libcolorsimplify.so")))
lib.cs_last_error.restype = C.c_char_p
vp = C.c_void_p
ctx = vp(); assert lib.cs_ctx_create(0, C.byref(ctx)) == 0
dev = torch.device("cuda:0")
n = int(sys.argv[1]); K = int(sys.argv[2]); variants = [int(v) for v in sys.argv[3].split(",")]
g = torch.Generator(device=dev); g.manual_seed(3)
planes = [torch.rand(n, device=dev, generator=g) * s + o for s, o in ((100, 0), (185, -90), (200, -105))]
idx = torch.randint(0, n, (K,), device=dev, generator=g)
Cn = torch.stack([p[idx] for p in planes], 1).double().contiguous()
d_lab = torch.empty(n, dtype=torch.uint8, device=dev)
d_sums = torch.zeros(K * 3, dtype=torch.float64, device=dev); d_cnt = torch.zeros(K, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
ref = None
for v in variants:
	for fl in (0, 1):
		flags = fl | (v << 8)
		def run():
			rc = lib.cs_lloyd_step_f32(ctx, vp(planes[0].data_ptr()), vp(planes[1].data_ptr()), vp(planes[2].data_ptr()),
			                           C.c_int64(n), vp(Cn.data_ptr()), K, vp(d_lab.data_ptr()), vp(d_sums.data_ptr()),
			                           vp(d_cnt.data_ptr()), None, C.c_double(31400.0), flags, vp(st))
			assert rc == 0, lib.cs_last_error()
		for _ in range(3): run()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		torch.cuda.synchronize(); e0.record()
		for _ in range(20): run()
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1) / 20
		chk = (int(d_lab.long().sum().item()), d_cnt.cpu().tolist(), d_sums.cpu())
		if fl == 1:
			if ref is None: ref = chk
			same = chk[0] == ref[0] and chk[1] == ref[1] and torch.allclose(chk[2], ref[2], rtol=1e-6)
		else:
			same = "-"
		print(f"n={n} K={K} v={v} fl={fl} ms={ms:.4f} hbm={13.0*n/ms/1e6/6549.8*100:.1f}% same_as_first={same}", flush=True)
