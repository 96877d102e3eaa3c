"""GPU probe for the Lloyd kernel: correctness against a numpy fp64 restatement on small
cases, then a variant sweep at 64 MP, K=16 (CUDA-event timing).  Development tool — the
judged numbers come from bench.py; results land in gpurun_out/lloyd_probe.json."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
lib = C.CDLL(str(ROOT / "image_segmenter_b200/_lib/libcolorsimplify.so"))
lib.cs_last_error.restype = C.c_char_p
vp = C.c_void_p


def chk(rc, what):
	if rc != 0:
		raise RuntimeError(f"{what}: rc={rc} {lib.cs_last_error().decode()}")


X2MAX = 100.0 ** 2 + 95.0 ** 2 + 105.0 ** 2 + 1.0
ctx = vp()
chk(lib.cs_ctx_create(0, C.byref(ctx)), "ctx")
dev = torch.device("cuda:0")


def step(planes, centers, K, flags, labels=True, inertia=False, fused=False):
	n = planes[0].numel()
	d_c = torch.as_tensor(centers, dtype=torch.float64, device=dev).contiguous()
	d_lab = torch.full((max(n, 1) + 3,), 77, dtype=torch.uint8, device=dev) if labels else None
	d_sums = torch.zeros(K * 3, dtype=torch.float64, device=dev)
	d_cnt = torch.zeros(K, dtype=torch.float64, device=dev)
	d_in = torch.zeros(1, dtype=torch.float64, device=dev) if inertia else None
	st = torch.cuda.current_stream().cuda_stream
	if fused:
		d_out = torch.zeros(K * 3, dtype=torch.float64, device=dev)
		d_stats = torch.zeros(4, dtype=torch.float64, device=dev)
		chk(lib.cs_lloyd_iter_f32(ctx, vp(planes[0].data_ptr()), vp(planes[1].data_ptr()), vp(planes[2].data_ptr()),
		                          C.c_int64(n), vp(d_c.data_ptr()), K, vp(d_lab.data_ptr()) if labels else None,
		                          vp(d_sums.data_ptr()), vp(d_cnt.data_ptr()), vp(d_out.data_ptr()),
		                          vp(d_stats.data_ptr()), C.c_double(X2MAX), flags, vp(st)), "iter")
		torch.cuda.synchronize()
		return (d_lab[:n].cpu().numpy() if labels else None, d_sums.cpu().numpy().reshape(K, 3), d_cnt.cpu().numpy(),
		        d_out.cpu().numpy().reshape(K, 3), d_stats.cpu().numpy())
	chk(lib.cs_lloyd_step_f32(ctx, vp(planes[0].data_ptr()), vp(planes[1].data_ptr()), vp(planes[2].data_ptr()),
	                          C.c_int64(n), vp(d_c.data_ptr()), K, vp(d_lab.data_ptr()) if labels else None,
	                          vp(d_sums.data_ptr()), vp(d_cnt.data_ptr()), vp(d_in.data_ptr()) if inertia else None,
	                          C.c_double(X2MAX), flags, vp(st)), "step")
	torch.cuda.synchronize()
	return (d_lab[:n].cpu().numpy() if labels else None, d_sums.cpu().numpy().reshape(K, 3), d_cnt.cpu().numpy(),
	        float(d_in.item()) if inertia else None, d_lab[n:].cpu().numpy() if labels else None)


def ref_step(X, Cn):
	X = X.astype(np.float64)
	d = np.zeros((len(X), Cn.shape[0]))
	for j in range(3):
		d += (X[:, j, None] - Cn[None, :, j]) ** 2
	lab = d.argmin(1)
	K = Cn.shape[0]
	sums = np.zeros((K, 3)); cnt = np.zeros(K)
	np.add.at(sums, lab, X); np.add.at(cnt, lab, 1.0)
	ds = np.sort(d, axis=1)
	gap = ds[:, 1] - ds[:, 0] if K > 1 else np.full(len(X), np.inf)
	return lab, sums, cnt, d.min(1).sum(), gap, (ds[:, 1] if K > 1 else np.zeros(len(X)))


out = {"correctness": [], "sweep": []}
rng = np.random.default_rng(0)
ok_all = True
for n in [1, 3, 4, 5, 1000, 2047, 2048, 2049, 100003, 1 << 20]:
	for K in [2, 5, 8, 16, 17, 64, 200, 256]:
		if n > 200000 and K > 64:
			continue
		X = np.stack([rng.uniform(0, 100, n), rng.uniform(-90, 95, n), rng.uniform(-105, 95, n)], 1).astype(np.float32)
		Cn = X[rng.choice(n, K, replace=(n < K))].astype(np.float64) + (rng.normal(0, 1e-3, (K, 3)) if n < K else 0)
		planes = [torch.from_numpy(np.ascontiguousarray(X[:, j])).to(dev) for j in range(3)]
		rl, rs, rc, ri, gap, dsec = ref_step(X, Cn)
		for flags in (0, 1):
			lab, sums, cnt, inert, guard = step(planes, Cn, K, flags, inertia=True)
			mism = np.nonzero(lab != rl)[0]
			# every mismatch must be a near tie under the documented fp32 bound
			xx = (X.astype(np.float64) ** 2).sum(1)
			bits = max(3, int(np.ceil(np.log2(max(K, 8)))))
			vtop = (np.sqrt(X2MAX) + np.sqrt((Cn ** 2).sum(1).max())) ** 2
			tau_int = np.full(len(X), 1.5 * 2 * 7 * 2.0 ** -24 * (vtop + 1024.0))
			tau_flt = 1.5 * 2.2 * (5 * 2.0 ** -24 * (xx + 2 * (Cn ** 2).sum(1).max()) + 2.0 ** -(23 - bits) * dsec)
			tau = tau_int if K <= 64 else tau_flt
			bad = [int(i) for i in mism if gap[i] > tau[i]]
			guard_ok = bool((guard == 77).all())
			if flags == 1:
				lab_ok = len(mism) == 0 or all(gap[i] < 1e-9 for i in mism)
			else:
				lab_ok = len(bad) == 0
			# sums must match the labels the kernel chose
			s2 = np.zeros((K, 3)); c2 = np.zeros(K)
			np.add.at(s2, lab, X.astype(np.float64)); np.add.at(c2, lab, 1.0)
			sums_ok = np.allclose(sums, s2, rtol=2e-6, atol=1e-3) and np.array_equal(cnt, c2)
			inert_ok = abs(inert - ri) <= 1e-4 * max(ri, 1.0) + 1e-2
			ok = lab_ok and sums_ok and guard_ok and inert_ok
			ok_all &= ok
			out["correctness"].append(dict(n=n, K=K, flags=flags, mismatches=len(mism), bad=len(bad), sums_ok=bool(sums_ok),
			                              guard_ok=guard_ok, inert_ok=bool(inert_ok), ok=bool(ok)))
			if not ok:
				print("FAIL", out["correctness"][-1], flush=True)
		# fused iteration == step + host finalize
		lab, sums, cnt, cnew, stats = step(planes, Cn, K, 1, fused=True)
		exp = np.where(cnt[:, None] > 0, sums * (1.0 / np.maximum(cnt, 1))[:, None], np.nan)
		nz = cnt > 0
		f_ok = np.array_equal(cnew[nz], exp[nz]) and stats[1] == (K - nz.sum())
		ok_all &= bool(f_ok)
		if not f_ok:
			print("FUSED FAIL", n, K, stats, flush=True)
print("correctness ok:", ok_all, flush=True)
out["correctness_ok"] = bool(ok_all)

# ---- variant sweep at 64 MP (and 8 MP), K = 16 ----
def sweep(n, K, variants, flagsets, iters=10):
	g = torch.Generator(device=dev); g.manual_seed(3)
	planes = [torch.rand(n, device=dev, generator=g) * s + o for s, o in ((100, 0), (185, -90), (200, -105))]
	idx = torch.randint(0, n, (K,), device=dev, generator=g)
	Cn = torch.stack([p[idx] for p in planes], 1).double().contiguous()
	d_lab = torch.empty(n, dtype=torch.uint8, device=dev)
	d_sums = torch.zeros(K * 3, dtype=torch.float64, device=dev); d_cnt = torch.zeros(K, dtype=torch.float64, device=dev)
	st = torch.cuda.current_stream().cuda_stream
	res = []
	for v in variants:
		for fl in flagsets:
			flags = fl | (v << 8)
			def run():
				chk(lib.cs_lloyd_step_f32(ctx, vp(planes[0].data_ptr()), vp(planes[1].data_ptr()), vp(planes[2].data_ptr()),
				                          C.c_int64(n), vp(Cn.data_ptr()), K, vp(d_lab.data_ptr()), vp(d_sums.data_ptr()),
				                          vp(d_cnt.data_ptr()), None, C.c_double(X2MAX), flags, vp(st)), "step")
			for _ in range(3):
				run()
			torch.cuda.synchronize()
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record()
			for _ in range(iters):
				run()
			e1.record(); torch.cuda.synchronize()
			ms = e0.elapsed_time(e1) / iters
			gbs = 13.0 * n / (ms * 1e-3) / 1e9
			r = dict(n=n, K=K, variant=v, flags=fl, ms=ms, mpix_s=n / ms / 1e3, gbs=gbs, frac_hbm=gbs / 6549.8,
			         cnt_sum=float(d_cnt.sum().item()))
			print(r, flush=True)
			res.append(r)
	return res

out["sweep"] += sweep(8192 * 8192, 16, range(0, 12), (0, 1))
out["sweep"] += sweep(3840 * 2160, 16, (0, 4), (0, 1))
out["sweep"] += sweep(8192 * 8192, 64, (0,), (0, 1), iters=5)
out["sweep"] += sweep(8192 * 8192, 32, (0,), (0, 1), iters=5)
out["sweep"] += sweep(8192 * 8192, 8, (0,), (0, 1), iters=5)
out["sweep"] += sweep(8192 * 8192, 256, (0,), (0,), iters=3)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out/lloyd_probe.json", "w"), indent=1)
print("done")
