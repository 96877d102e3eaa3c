#!/usr/bin/env python
"""bench.py — LAB k-means MPix/s per Lloyd iteration (assign + update, labels written), K=16,
8192 x 8192 pixels per GPU, on N B200s; fraction of HBM roofline; the reference's scikit-learn CPU
kernel timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--k 16] [--exact]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one Lloyd iteration over every rank's resident shard (3 fp32 LAB planes in HBM, 805 MB
per GPU — larger than the 126 MB L2, so nothing is served from cache between iterations).  Rank 0
prints ONE JSON line.  `value` is device-timed with inputs resident; `e2e` is the same metric
through the host-buffer C-ABI call (pinned host RGBA in, labels + centres out, copies inside the
timed region).  The oracle / scikit-learn are touched only by the cpu_baseline leg and
`--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
	sys.path.insert(0, str(ROOT))

H = W = 8192
BLOCK_ROWS = 1024
METRIC = "LAB k-means MPix/s per iteration (k=16, 64MP)"
UNIT = "MPix/s"
BYTES_PER_PX = 13.0  # 12 B read (L, a, b fp32) + 1 B written (u8 label): SURVEY.md §8d


def parse():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=200)
	ap.add_argument("--warmup", type=int, default=10)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--k", type=int, default=16)
	ap.add_argument("--exact", action="store_true", help="EXACT_TIES mode (fp64 re-evaluation of near ties)")
	ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
	                help="weak: 64 MP per GPU (default); strong: one 64 MP image row-sharded over the GPUs")
	ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p"])
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-e2e", action="store_true")
	return ap.parse_args()


def synth_block(seed: int, block: int, rows: int = BLOCK_ROWS) -> np.ndarray:
	"""rows x W RGBA8 block of the seeded uniform-random image (alpha 255); block-wise seeds so that
	every rank builds its rows without the others'."""
	rng = np.random.default_rng([seed, block])
	out = np.empty((rows, W, 4), dtype=np.uint8)
	out[:, :, :3] = rng.integers(0, 256, (rows, W, 3), dtype=np.uint8)
	out[:, :, 3] = 255
	return out


def initial_centers(k: int) -> np.ndarray:
	"""K distinct seeded sRGB colours in LAB — every colour occurs in the uniform-random image, so these
	are data points; identical on every rank and for the CPU baseline."""
	from image_segmenter_b200 import _colorspace as cspace

	rng = np.random.default_rng(16)
	return cspace.rgb2lab_small(rng.integers(0, 256, (k, 3), dtype=np.uint8))


class ClockSampler:
	"""SM clock + throttle reasons via NVML while the timed region runs."""

	def __init__(self, index: int):
		self.samples, self.reasons, self.max_mhz = [], set(), None
		self._stop = threading.Event()
		self._thr = None
		try:
			import pynvml

			pynvml.nvmlInit()
			self.nv = pynvml
			self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
			self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
		except Exception:
			self.nv = None

	def sample(self):
		if not self.nv:
			return
		nv = self.nv
		try:
			self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
			r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
				else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
			names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
			         0x80: "hw_power_brake_slowdown"}
			for bit, name in names.items():
				if r & bit:
					self.reasons.add(name)
		except Exception:
			pass

	def start(self):
		def loop():
			while not self._stop.is_set():
				self.sample()
				time.sleep(0.0005)

		self._thr = threading.Thread(target=loop, daemon=True)
		self._thr.start()

	def stop(self):
		self._stop.set()
		if self._thr:
			self._thr.join()
		return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
		        "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(k: int, C0: np.ndarray, budget_px: int = 1 << 24, iters: int = 3):
	"""scikit-learn's lloyd_iter_chunked_dense (the reference's KMeans kernel), fp64, all host threads, on a
	16 MP sample of the same synthetic distribution x 3 iterations."""
	from oracle import cpu_baseline as cb

	X = cb.make_lab_sample(budget_px, 3)
	secs, kind, threads, _ = cb.time_lloyd_iterations(X, C0, iters)
	val = budget_px * iters / secs / 1e6
	return {"value": round(val, 2), "unit": UNIT, "cores": threads, "kind": kind,
	        "sample": f"{budget_px} px (16 MP of the same uniform-random LAB distribution) x {iters} Lloyd iterations, "
	                  f"sklearn lloyd_iter_chunked_dense fp64, {threads} OpenMP threads of {os.cpu_count()} host cpus"}


def run_reference(args):
	"""--impl reference: the reference's own CPU Lloyd kernel on the host cores (rank 0 only)."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	C0 = initial_centers(args.k)
	from oracle import cpu_baseline as cb

	px = 1 << 23  # 8 MP per step keeps `--steps K` bounded (about 0.1-0.2 s per step on 8+ cores)
	X = cb.make_lab_sample(px, 3)
	cb.time_lloyd_iterations(X, C0, max(1, args.warmup))
	secs, kind, threads, _ = cb.time_lloyd_iterations(X, C0, args.steps)
	val = px * args.steps / secs / 1e6
	sample = (f"each step = one Lloyd iteration over an 8 MP sample of the 64 MP workload, sklearn "
	          f"lloyd_iter_chunked_dense fp64, {threads} OpenMP threads of {os.cpu_count()} host cpus")
	line = {"impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus,
	        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs / args.steps * 1e3, 3),
	        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
	        "config": workload_config(args, 1),
	        "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
	        "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	        "gpu_launches": 0}
	print(json.dumps(line), flush=True)


def workload_config(args, world):
	return {"workload": f"LAB k-means K={args.k}, one Lloyd iteration (assign+update, u8 labels written) over "
	                    f"{H}x{W} px per GPU, seeded uniform-random sRGB (seed 3) converted to fp32 CIELAB planes",
	        "k": args.k, "pixels_per_gpu": H * W if args.scaling == "weak" else H * W // world,
	        "image": f"{H * world if args.scaling == 'weak' else H}x{W}", "sharding": "contiguous row blocks",
	        "exchange": "none (1 GPU)" if world == 1 else (
	            "NCCL all_reduce of 4K doubles + finalize kernel" if args.exchange == "nccl" else
	            "fused in the Lloyd kernel: partials stored to peer mailboxes over NVLink P2P, reduced in rank order"),
	        "label_mode": "exact_ties" if args.exact else "fast (fp32 keys; mismatches only within the documented near-tie bound)",
	        "l2": "inputs larger than L2 (805 MB of planes per GPU vs 126 MB)" if args.scaling == "weak" or world == 1
	              else "shard may fit L2 at N>=8 (strong scaling)"}


def main():
	args = parse()
	if args.impl == "reference":
		return run_reference(args)

	import torch
	import torch.distributed as dist

	from image_segmenter_b200 import _ffi
	from image_segmenter_b200.engine import get_engine
	from image_segmenter_b200.sharded import make_gpu_lloyd, shard_rows

	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
	torch.cuda.set_device(local)
	if world > 1:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local))
	eng = get_engine(local)
	K = args.k

	# ---- this rank's shard: RGBA8 blocks -> device -> fp32 LAB planes (setup, untimed) ----
	if args.scaling == "weak":
		blocks = [(rank * (H // BLOCK_ROWS) + b) for b in range(H // BLOCK_ROWS)]
	else:
		r0, r1 = shard_rows(H, world, rank)
		assert r0 % BLOCK_ROWS == 0 and r1 % BLOCK_ROWS == 0, "strong scaling needs N | 8"
		blocks = list(range(r0 // BLOCK_ROWS, r1 // BLOCK_ROWS))
	n_local = len(blocks) * BLOCK_ROWS * W
	host_rgba = torch.empty((n_local, 4), dtype=torch.uint8).pin_memory()
	hv = host_rgba.numpy()
	for i, b in enumerate(blocks):
		hv[i * BLOCK_ROWS * W:(i + 1) * BLOCK_ROWS * W] = synth_block(3, b).reshape(-1, 4)
	d_rgba = host_rgba.to(eng.dev, non_blocking=False)
	planes = eng.rgba_to_lab(d_rgba)
	del d_rgba
	labels = torch.empty(n_local, dtype=torch.uint8, device=eng.dev)
	C0 = initial_centers(K)
	exchange = args.exchange if args.exchange != "auto" else ("p2p" if world > 1 else "none")
	drv = make_gpu_lloyd(eng, planes, n_local, K, labels=labels, exact=args.exact, exchange=args.exchange)
	drv.set_centers(C0)
	launches_per_step = 2 if (world > 1 and exchange == "nccl") else 1

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	# ---- warm-up: clocks + W untimed steps ----
	for _ in range(50):
		drv.iterate()
	drv.set_centers(C0)
	for _ in range(max(3, args.warmup)):
		drv.iterate()
	barrier()

	# ---- timed region: exactly K steps, CUDA events on the launching (current) stream ----
	sampler = ClockSampler(local)
	# (one event on each side only: an event between two launches would keep the next launch's prologue
	# from starting under the previous launch's tail — CS_LLOYD_CHAINED)
	ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
	sampler.start()
	ev[0].record()
	for i in range(args.steps):
		drv.iterate()
	ev[1].record()
	sampler.sample()
	barrier()
	clocks = sampler.stop()
	total_ms = ev[0].elapsed_time(ev[1])
	per_step = np.array([total_ms / args.steps])
	t = torch.tensor([total_ms], dtype=torch.float64, device=eng.dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	total_ms = float(t.item())
	final_stats = drv.stats.cpu().numpy()
	counts_total = float(drv.acc[3 * K:].sum().item())

	value = n_local * world * args.steps / (total_ms * 1e-3) / 1e6
	peaks = {}
	try:
		peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
	except Exception:
		pass
	peak = float(peaks.get("hbm_gbs", 6650.0))
	kern_ms = float(np.mean(per_step))  # the step is one kernel (fused M-step tail) at N=1
	achieved = BYTES_PER_PX * n_local / (kern_ms * 1e-3) / 1e9
	traffic = None
	try:
		prof = json.loads((ROOT / "profiles" / "lloyd_traffic.json").read_text())
		traffic = prof.get(f"k{K}_{'exact' if args.exact else 'fast'}_dram_bytes_per_launch")
	except Exception:
		pass
	roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak,
	            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
	            "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
	            "kernel": "lloyd_kernel (assign + update + fused M-step tail)",
	            "kernel_ms": round(kern_ms, 5), "algorithmic_bytes_per_launch": BYTES_PER_PX * n_local}

	line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
	        "warmup": max(3, args.warmup), "ms_per_step": round(total_ms / args.steps, 5), "higher_is_better": True,
	        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
	        "config": workload_config(args, world), "roofline": roofline, "clocks": clocks,
	        "gpu_launches": launches_per_step * args.steps,
	        "check": {"count_sum_last_step": counts_total, "expected": float(n_local * world),
	                  "shift2_last_step": float(final_stats[0])}}

	# ---- N > 1: every rank's own speed on its shard WITHOUT the exchange (50 fused single-GPU iterations),
	# so that the gap between the per-step time above and a single GPU can be attributed: the exchange makes
	# all ranks wait for the slowest one in every iteration ----
	if world > 1:
		from image_segmenter_b200.engine import KMeansGPU
		km = KMeansGPU(eng, "f32", n_local, planes=planes, exact=args.exact)
		ca, cb = torch.from_numpy(C0.copy()).to(eng.dev), torch.zeros((K, 3), dtype=torch.float64, device=eng.dev)
		ssum, scnt = torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev)
		sst = torch.zeros(4, dtype=torch.float64, device=eng.dev)
		sctl = torch.tensor([0.0, 0.0, -1.0, 0.0], dtype=torch.float64, device=eng.dev)
		km._run(ca, cb, K, ssum, scnt, sst, 10, sctl)
		barrier()
		s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		s0.record()
		km._run(ca, cb, K, ssum, scnt, sst, 50, sctl)
		s1.record()
		torch.cuda.synchronize()
		mine = torch.tensor([s0.elapsed_time(s1) / 50, float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device=eng.dev)
		allr = [torch.zeros_like(mine) for _ in range(world)]
		dist.all_gather(allr, mine)
		line["per_rank_standalone"] = {"ms_per_iteration_no_labels": [round(float(t[0]), 5) for t in allr],
		                                "sm_mhz_in_timed_region": [float(t[1]) for t in allr],
		                                "what": "each rank alone on its shard, fused iterations without exchange and without the label store"}
		barrier()

	# ---- the other label mode, for information (same shard, 50 steps, CUDA events) ----
	if not args.exact:
		drv_x = make_gpu_lloyd(eng, planes, n_local, K, labels=labels, exact=True, exchange=args.exchange)
		drv_x.set_centers(C0)
		for _ in range(5):
			drv_x.iterate()
		barrier()
		x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		x0.record()
		for _ in range(50):
			drv_x.iterate()
		x1.record()
		barrier()
		tx = torch.tensor([x0.elapsed_time(x1)], dtype=torch.float64, device=eng.dev)
		if world > 1:
			dist.all_reduce(tx, op=dist.ReduceOp.MAX)
		ms_x = float(tx.item()) / 50
		line["exact_ties_mode"] = {"value": round(n_local * world / (ms_x * 1e-3) / 1e6, 1), "unit": UNIT,
		                           "ms_per_step": round(ms_x, 5),
		                           "roofline_frac": round(BYTES_PER_PX * n_local / (ms_x * 1e-3) / 1e9 / peak, 4),
		                           "what": "same step with CS_LLOYD_EXACT_TIES: labels equal the fp64 first-minimum"}

	# ---- e2e: host buffers through the C ABI (H2D + LAB + 20 iterations + D2H of labels/centres) ----
	if not args.no_e2e:
		e2e_iters = 20
		lut = np.ascontiguousarray(__import__("image_segmenter_b200._colorspace", fromlist=["x"]).linear_lut256())
		h_labels = torch.empty(n_local, dtype=torch.uint8).pin_memory()
		reps = 3
		import ctypes as C

		times = []
		for rep in range(reps + 1):
			cen = np.ascontiguousarray(C0, dtype=np.float64).copy()
			nit, inert = C.c_int(0), C.c_double(0.0)
			barrier()
			t0 = time.perf_counter()
			if world == 1:
				_ffi.check(eng.ctx.lib.cs_host_lab_kmeans(eng.ctx.handle, host_rgba.data_ptr(), n_local, lut.ctypes.data,
				                                          cen.ctypes.data, K, e2e_iters, 0.0, 1 if args.exact else 0,
				                                          h_labels.data_ptr(), C.byref(nit), C.byref(inert)),
				           "cs_host_lab_kmeans")
			else:
				d_in = host_rgba.to(eng.dev, non_blocking=True)
				pl = eng.rgba_to_lab(d_in)
				lab2 = torch.empty(n_local, dtype=torch.uint8, device=eng.dev)
				d2 = make_gpu_lloyd(eng, pl, n_local, K, labels=lab2, exact=args.exact, exchange=args.exchange)
				d2.set_centers(C0)
				for _ in range(e2e_iters):
					d2.iterate()
				h_labels.copy_(lab2, non_blocking=True)
				cen = d2.c[d2.cur].cpu().numpy()
			barrier()
			if rep:  # first call pays the allocation of the library's staging buffer
				times.append(time.perf_counter() - t0)
		tt = torch.tensor([min(times)], dtype=torch.float64, device=eng.dev)
		if world > 1:
			dist.all_reduce(tt, op=dist.ReduceOp.MAX)
		e2e_val = n_local * world * e2e_iters / float(tt.item()) / 1e6
		line["e2e"] = {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": n_local * 4 + 2048 + K * 24,
		               "d2h_bytes_per_step": n_local + K * 24 + 8,
		               "what": f"one host-buffer call = upload RGBA8 + LAB conversion + {e2e_iters} Lloyd iterations + "
		                       f"final E-step + download labels/centres; value = pixels x {e2e_iters} / wall time",
		               "ms_per_call": round(float(tt.item()) * 1e3, 3)}

	if rank == 0 and not args.no_cpu_baseline:
		line["cpu_baseline"] = cpu_baseline(K, C0)
	if rank == 0:
		print(json.dumps(line), flush=True)
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()


if __name__ == "__main__":
	main()
