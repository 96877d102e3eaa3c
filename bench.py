#!/usr/bin/env python
"""bench.py — LAB k-means MPix/s per Lloyd iteration (assign + update, labels written), K=16,
8192 x 8192 pixels per GPU, on N B200s; fraction of HBM roofline; the reference's scikit-learn CPU
kernel timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--k 16] [--fast]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one Lloyd iteration over every rank's resident shard (3 fp32 LAB planes in HBM, 805 MB
per GPU — larger than the 126 MB L2, so nothing is served from cache between iterations).  Rank 0
prints ONE JSON line.

  value      device-timed, inputs resident, in the label mode the PRODUCT runs (CS_LLOYD_EXACT_TIES:
             labels equal the fp64 first minimum, i.e. the oracle's); `fast_mode` holds the fp32-key
             mode for information (`--fast` swaps the two).
  e2e        the same metric through the public drop-in API with HOST buffers:
             simplify_colors_perceptual_fast(rgba, K, fit="full", init_centers=..., max_iter=20) — upload,
             LAB conversion, 20 Lloyd iterations, nearest-centre remap, download of the RGBA result, all
             inside the timed call (at N > 1 every rank calls it on its rows with process_group=WORLD).
  configs    the other BASELINE.json configs, each with its own roofline block (K=64 at 64 MP, one 64 MP
             image row-sharded over the N GPUs, the 1024 x 1080p batch partitioned over the N GPUs, the
             16 MP integer paths with bit_exact booleans).
  check      at N > 1: centres bit-equal across ranks, fused P2P exchange vs NCCL all_reduce path, sharded
             vs unsharded run of a sub-image.
The oracle / scikit-learn are touched only by the cpu_baseline leg and `--impl reference`.
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
E2E_ITERS = 20


def parse():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=200)
	ap.add_argument("--warmup", type=int, default=10)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--k", type=int, default=16)
	ap.add_argument("--fast", action="store_true", help="headline in the fp32-key label mode (default: EXACT_TIES, the product's)")
	ap.add_argument("--exact", action="store_true", help="(default; kept for compatibility)")
	ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
	                help="weak: 64 MP per GPU (default); strong: one 64 MP image row-sharded over the GPUs")
	ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p"])
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-e2e", action="store_true")
	ap.add_argument("--no-configs", action="store_true", help="skip the legs for the other BASELINE configs")
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

		self.samples.clear()
		self._stop.clear()
		self._thr = threading.Thread(target=loop, daemon=True)
		self._thr.start()

	def stop(self):
		self._stop.set()
		if self._thr:
			self._thr.join()
		return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
		        "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_cores() -> int:
	"""The cores this process may run on — NOT the OpenMP default, which torchrun pins to 1
	(OMP_NUM_THREADS=1) in every worker it spawns."""
	try:
		return len(os.sched_getaffinity(0))
	except Exception:
		return os.cpu_count() or 1


def cpu_baseline(k: int, C0: np.ndarray, budget_px: int = 1 << 24, iters: int = 3):
	"""scikit-learn's lloyd_iter_chunked_dense (the reference's KMeans kernel), fp64, all host cores, on a
	16 MP sample of the same synthetic distribution x 3 iterations."""
	from oracle import cpu_baseline as cb

	X = cb.make_lab_sample(budget_px, 3)
	secs, kind, threads, _ = cb.time_lloyd_iterations(X, C0, iters, threads=host_cores())
	val = budget_px * iters / secs / 1e6
	return {"value": round(val, 2), "unit": UNIT, "cores": threads, "kind": kind,
	        "sample": f"{budget_px} px (16 MP of the same uniform-random LAB distribution) x {iters} Lloyd iterations, "
	                  f"sklearn lloyd_iter_chunked_dense fp64, {threads} OpenMP threads of {os.cpu_count()} host cpus"}


def run_reference(args):
	"""--impl reference: the reference's own CPU Lloyd kernel on ALL host cores (rank 0 only; the other ranks
	of a torchrun launch exit at once), same `config` as the repo arm at the same N."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	world = int(os.environ.get("WORLD_SIZE", str(max(1, args.gpus))))
	C0 = initial_centers(args.k)
	from oracle import cpu_baseline as cb

	threads = host_cores()
	px = 1 << 23  # 8 MP per step keeps `--steps K` bounded (about 0.03-0.2 s per step on 8+ cores)
	X = cb.make_lab_sample(px, 3)
	cb.time_lloyd_iterations(X, C0, max(1, args.warmup), threads=threads)
	secs, kind, threads, _ = cb.time_lloyd_iterations(X, C0, args.steps, threads=threads)
	val = px * args.steps / secs / 1e6
	sample = (f"each step = one Lloyd iteration over an 8 MP sample of the workload, sklearn "
	          f"lloyd_iter_chunked_dense fp64, {threads} OpenMP threads of {os.cpu_count()} host cpus")
	line = {"impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": world,
	        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(secs / args.steps * 1e3, 3),
	        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
	        "config": workload_config(args, world),
	        "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
	        "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	        "gpu_launches": 0}
	print(json.dumps(line), flush=True)


def workload_config(args, world):
	"""Identical for the repo arm and the reference arm at the same N (the driver compares them)."""
	return {"workload": f"LAB k-means K={args.k}, one Lloyd iteration (assign+update, u8 labels written) over "
	                    f"{H}x{W} px per GPU, seeded uniform-random sRGB (seed 3) converted to fp32 CIELAB planes",
	        "k": args.k, "pixels_per_gpu": H * W if args.scaling == "weak" else H * W // world,
	        "image": f"{H * world if args.scaling == 'weak' else H}x{W}", "sharding": "contiguous row blocks",
	        "exchange": "none (1 GPU)" if world == 1 else (
	            "NCCL all_reduce of 4K doubles + finalize kernel" if args.exchange == "nccl" else
	            "fused in the Lloyd kernel: partials stored to peer mailboxes over NVLink P2P, reduced in rank order"),
	        "label_mode": "fast (fp32 keys; mismatches only within the documented near-tie bound)" if args.fast else
	                      "exact_ties (labels == fp64 first minimum, the oracle's and the product's mode)",
	        "l2": "inputs larger than L2 (805 MB of planes per GPU vs 126 MB)" if args.scaling == "weak" or world == 1
	              else "shard may fit L2 at N>=8 (strong scaling)"}


def main():
	args = parse()
	if args.impl == "reference":
		return run_reference(args)

	import torch
	import torch.distributed as dist

	from image_segmenter_b200 import _ffi
	from image_segmenter_b200 import color_simplify as cs
	from image_segmenter_b200.engine import KMeansGPU, get_engine
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
	exact = not args.fast
	peaks = {}
	try:
		peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
	except Exception:
		pass
	peak = float(peaks.get("hbm_gbs", 6650.0))
	peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"

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

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	def max_over_ranks(ms: float) -> float:
		if world == 1:
			return ms
		t = torch.tensor([ms], dtype=torch.float64, device=eng.dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return float(t.item())

	def timed_steps(drv, steps: int, lead: int = 3, sampler=None):
		"""K steps between two CUDA events on the launching stream.  Everything that takes host time is done
		BEFORE the barrier; after it, `lead` untimed iterations are queued and the first event is recorded with
		no host synchronisation in between — with the in-kernel exchange the lead iterations lock-step the
		GPUs, so start skew between the processes is not charged to the timed steps.  One event on each side
		only: an event between two launches would keep a chained launch's prologue from starting under its
		predecessor's tail."""
		ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
		barrier()
		if sampler is not None:
			sampler.start()
		for _ in range(lead):
			drv.iterate()
		ev[0].record()
		for _ in range(steps):
			drv.iterate()
		ev[1].record()
		if sampler is not None:
			sampler.sample()
		barrier()
		return max_over_ranks(ev[0].elapsed_time(ev[1]))

	def roofline(ms: float, n_px: int, bytes_px: float, kernel: str, traffic=None, note=None):
		ach = bytes_px * n_px / (ms * 1e-3) / 1e9
		r = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
		     "frac": round(ach / peak, 4), "traffic": traffic, "kernel": kernel, "kernel_ms": round(ms, 5),
		     "algorithmic_bytes_per_launch": bytes_px * n_px}
		if note:
			r["note"] = note
		return r

	# ---- headline: warm-up (clocks + W untimed steps), then exactly K timed steps ----
	drv = make_gpu_lloyd(eng, planes, n_local, K, labels=labels, exact=exact, exchange=args.exchange)
	drv.set_centers(C0)
	for _ in range(50):
		drv.iterate()
	drv.set_centers(C0)
	for _ in range(max(3, args.warmup)):
		drv.iterate()
	sampler = ClockSampler(local)  # NVML initialisation happens here, before the barrier
	total_ms = timed_steps(drv, args.steps, sampler=sampler)
	clocks = sampler.stop()
	final_stats = drv.stats.cpu().numpy()
	counts_total = float(drv.acc[3 * K:].sum().item())
	# exact labels, 9 <= K <= 64: the grid-filtered assignment = a ~10 us candidate-table kernel + the Lloyd kernel
	grid_path = exact and 9 <= K <= 64 and n_local >= (10_000_000 if K <= 16 else 1 << 22 if K <= 32 else 1 << 21)
	launches_per_step = (2 if (world > 1 and exchange == "nccl") else 1) + (1 if grid_path else 0)
	kern_ms = total_ms / args.steps  # the fused Lloyd kernel (+ its table-build kernel on the grid path)
	value = n_local * world * args.steps / (total_ms * 1e-3) / 1e6
	traffic = None
	try:
		prof = json.loads((ROOT / "profiles" / "lloyd_traffic.json").read_text())
		traffic = prof.get(f"k{K}_{'exact' if exact else 'fast'}_dram_bytes_per_launch")
	except Exception:
		pass
	cfg = workload_config(args, world)
	line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
	        "warmup": max(3, args.warmup), "ms_per_step": round(total_ms / args.steps, 5), "higher_is_better": True,
	        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
	        "config": cfg,
	        "roofline": roofline(kern_ms, n_local, BYTES_PER_PX,
	                             "lloyd_kernel<GRID> (grid-filtered assign + update + fused M-step tail); kernel_ms includes the "
	                             "grid_build_kernel that precedes every launch" if grid_path else
	                             "lloyd_kernel (assign + update + fused M-step tail)", traffic),
	        "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
	        "check": {"count_sum_last_step": counts_total, "expected": float(n_local * world),
	                  "shift2_last_step": float(final_stats[0])}}

	# ---- the other label mode, for information (same shard, 50 steps) ----
	drv_o = make_gpu_lloyd(eng, planes, n_local, K, labels=labels, exact=not exact, exchange=args.exchange)
	drv_o.set_centers(C0)
	for _ in range(5):
		drv_o.iterate()
	ms_o = timed_steps(drv_o, 50) / 50
	line["fast_mode" if exact else "exact_ties_mode"] = {
		"value": round(n_local * world / (ms_o * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms_o, 5),
		"roofline_frac": round(BYTES_PER_PX * n_local / (ms_o * 1e-3) / 1e9 / peak, 4),
		"what": ("same step without CS_LLOYD_EXACT_TIES: fp32 keys, a pixel within the fp32 error bound of a tie may take "
		         "either centre") if exact else "same step with CS_LLOYD_EXACT_TIES: labels equal the fp64 first-minimum"}
	del drv_o

	# ---- N > 1: parity of the multi-GPU path, and every rank's own speed without the exchange ----
	if world > 1:
		line["check"].update(multi_gpu_parity(eng, planes, n_local, K, C0, drv, world, rank))
		km = KMeansGPU(eng, "f32", n_local, planes=planes, exact=exact)
		ca, cb_ = torch.from_numpy(C0.copy()).to(eng.dev), torch.zeros((K, 3), dtype=torch.float64, device=eng.dev)
		ssum, scnt = torch.zeros((K, 3), dtype=torch.float64, device=eng.dev), torch.zeros(K, dtype=torch.float64, device=eng.dev)
		sst = torch.zeros(4, dtype=torch.float64, device=eng.dev)
		sctl = torch.tensor([0.0, 0.0, -1.0, 0.0], dtype=torch.float64, device=eng.dev)
		km._run(ca, cb_, K, ssum, scnt, sst, 10, sctl)
		barrier()
		s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		s0.record()
		km._run(ca, cb_, K, ssum, scnt, sst, 50, sctl)
		s1.record()
		torch.cuda.synchronize()
		mine = torch.tensor([s0.elapsed_time(s1) / 50, float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device=eng.dev)
		allr = [torch.zeros_like(mine) for _ in range(world)]
		dist.all_gather(allr, mine)
		line["per_rank_standalone"] = {"ms_per_iteration_no_labels": [round(float(t[0]), 5) for t in allr],
		                                "sm_mhz_in_timed_region": [float(t[1]) for t in allr],
		                                "what": "each rank alone on its shard, fused iterations without exchange and without the label store"}
		barrier()

	# ---- the other BASELINE configs (each with its own roofline block) ----
	if not args.no_configs:
		line["configs"] = other_configs(args, eng, planes, n_local, labels, world, rank, timed_steps, roofline, barrier,
		                                max_over_ranks)

	# ---- e2e: host buffers through the public drop-in API ----
	if not args.no_e2e:
		del drv
		rows = n_local // W
		img = host_rgba.numpy().reshape(rows, W, 4)  # page-locked (torch pin_memory): stated in `what`
		group = dist.group.WORLD if world > 1 else None
		times = []
		for rep in range(3):
			barrier()
			t0 = time.perf_counter()
			out_img, pal = cs.simplify_colors_perceptual_fast(img, K, True, fit="full", init_centers=C0, max_iter=E2E_ITERS,
			                                                  tol=-1.0, process_group=group)
			torch.cuda.synchronize()
			dt = time.perf_counter() - t0
			barrier()
			if rep:  # the first call pays allocations (torch caching allocators, library scratch)
				times.append(dt)
			ok_shape = out_img.shape == img.shape and pal.shape == (K, 3)
			del out_img
		sec = max_over_ranks(min(times) * 1e3) * 1e-3
		e2e_val = n_local * world * E2E_ITERS / sec / 1e6
		line["e2e"] = {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": n_local * 4 + 2048 + K * 24 + K * 3,
		               "d2h_bytes_per_step": n_local * 4 + K * 24 + 64,
		               "what": f"one call of the public API simplify_colors_perceptual_fast(rgba, {K}, fit='full', init_centers=C0, "
		                       f"max_iter={E2E_ITERS}, tol=-1) on a page-locked HOST array: chunked upload overlapped with the LAB "
		                       f"conversion, {E2E_ITERS} Lloyd iterations (exact labels), nearest-centre remap, download of the RGBA "
		                       f"result; value = pixels x {E2E_ITERS} / wall time of the call (max over ranks)",
		               "ms_per_call": round(sec * 1e3, 3), "result_ok": bool(ok_shape)}
		if world == 1:
			# the same call on an ordinary pageable NumPy array (what a caller who does not register its buffer pays)
			# — the library stages it through page-locked buffers with its own copy threads, cs_host_upload)
			pageable = np.array(img, copy=True)
			pt = []
			for rep in range(3):
				t0 = time.perf_counter()
				o2, _ = cs.simplify_colors_perceptual_fast(pageable, K, True, fit="full", init_centers=C0, max_iter=E2E_ITERS, tol=-1.0)
				torch.cuda.synchronize()
				if rep:  # as above: the first call pays allocations (staging ring, copy threads)
					pt.append(time.perf_counter() - t0)
				del o2
			dt = min(pt)
			line["e2e"]["pageable_input"] = {"value": round(n_local * E2E_ITERS / dt / 1e6, 1), "ms_per_call": round(dt * 1e3, 3),
			                                 "what": "same call, ordinary (pageable) NumPy array: threaded staged upload (cs_host_upload)"}
			del pageable

	if rank == 0 and world == 1 and not args.no_cpu_baseline:
		line["cpu_baseline"] = cpu_baseline(K, C0)
	if rank == 0:
		print(json.dumps(line), flush=True)
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()


def multi_gpu_parity(eng, planes, n_local, K, C0, drv, world, rank):
	"""check fields at N > 1 (reference semantics: ONE KMeans over the whole image, identical centres —
	app/processing/color_simplify.py:669-675 via sklearn/cluster/_kmeans.py:705-738)."""
	import torch
	import torch.distributed as dist

	from image_segmenter_b200.engine import KMeansGPU
	from image_segmenter_b200.sharded import make_gpu_lloyd

	out = {}
	# (1) the centres every rank holds after the timed steps are bit-identical
	c = drv.c[drv.cur].clone()
	g = [torch.zeros_like(c) for _ in range(world)]
	dist.all_gather(g, c)
	out["centres_bit_equal_across_ranks"] = bool(all(torch.equal(g[0], x) for x in g))
	# (2) fused P2P exchange vs NCCL all_reduce path, 6 iterations from C0 on the full shards
	res = {}
	for ex in ("p2p", "nccl"):
		d2 = make_gpu_lloyd(eng, planes, n_local, K, exact=True, exchange=ex)
		r = d2.run(C0, 6, -1.0)
		res[ex] = r.centers
	out["p2p_vs_nccl_max_abs"] = float(np.abs(res["p2p"] - res["nccl"]).max())
	# (3) sharded vs UNSHARDED: a sub-image made of the first 1 MP of every rank's shard; rank 0 gathers the
	# pieces and runs the plain single-GPU loop on the concatenation
	m = 1 << 20
	sub = planes[:, :m].contiguous()
	d3 = make_gpu_lloyd(eng, sub, m, K, exact=True, exchange="p2p")
	r3 = d3.run(C0, 6, -1.0)
	pieces = [torch.zeros_like(sub) for _ in range(world)]
	dist.all_gather(pieces, sub)
	diff = torch.zeros(1, dtype=torch.float64, device=eng.dev)
	if rank == 0:
		whole = torch.cat(pieces, dim=1).contiguous()
		km = KMeansGPU(eng, "f32", m * world, planes=whole, exact=True)
		ref = km.fit_centers(C0, max_iter=6, tol=-1.0)
		diff[0] = float(np.abs(ref - r3.centers).max())
	dist.broadcast(diff, 0)
	out["sharded_vs_unsharded_max_abs"] = float(diff.item())
	out["parity_ok"] = bool(out["centres_bit_equal_across_ranks"] and out["p2p_vs_nccl_max_abs"] <= 1e-9
	                        and out["sharded_vs_unsharded_max_abs"] <= 1e-6)
	out["parity_what"] = ("6 Lloyd iterations from the same centres: (2) on the full shards, fused P2P exchange vs step + NCCL "
	                      "all_reduce + finalize; (3) first 1 MP of every shard, sharded run vs one GPU on the concatenation "
	                      "(per-CTA fp32 slot sums: the pixel->CTA split differs, hence not bit-equal)")
	return out


def other_configs(args, eng, planes, n_local, labels, world, rank, timed_steps, roofline, barrier, max_over_ranks):
	import torch

	from image_segmenter_b200 import _ffi
	from image_segmenter_b200 import color_simplify as cs
	from image_segmenter_b200.batch import partition_batch
	from image_segmenter_b200.sharded import make_gpu_lloyd

	out = {}
	# ---- config 3's own K: LAB k-means K=64 at 64 MP per GPU (exact labels) ----
	K64 = 64
	C64 = initial_centers(K64)
	d64 = make_gpu_lloyd(eng, planes, n_local, K64, labels=labels, exact=True, exchange=args.exchange)
	d64.set_centers(C64)
	for _ in range(5):
		d64.iterate()
	ms = timed_steps(d64, 20) / 20
	out["c3_k64_64mp_per_gpu"] = {
		"value": round(n_local * world / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms, 5), "label_mode": "exact_ties",
		"roofline": roofline(ms, n_local, BYTES_PER_PX, "lloyd_kernel<64, GRID> + grid_build_kernel (grid-filtered exact assignment)",
		                     note="bound by shared-memory wavefronts (four divergent 16-byte centre gathers per pixel), not by HBM; "
		                          "the exact full walk takes 1.26 ms")}
	del d64
	# ---- config 3 as specified: ONE 64 MP image row-sharded over the N GPUs (strong scaling) ----
	if world > 1 and args.scaling == "weak":
		n_s = (H * W // world) & ~3
		for kk, cc in ((args.k, initial_centers(args.k)), (K64, C64)):
			ds = make_gpu_lloyd(eng, planes[:, :n_s], n_s, kk, labels=labels[:n_s], exact=True, exchange=args.exchange)
			ds.set_centers(cc)
			for _ in range(10):
				ds.iterate()
			ms = timed_steps(ds, 50) / 50
			out[f"c3_one_64mp_image_sharded_k{kk}"] = {
				"value": round(n_s * world / (ms * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(ms, 5), "scaling": "strong",
				"pixels_per_gpu": n_s, "label_mode": "exact_ties",
				"roofline": roofline(ms, n_s, BYTES_PER_PX, "lloyd_kernel<%d%s> + fused exchange" % (
					kk, ", GRID" if n_s >= (10_000_000 if kk <= 16 else 1 << 22 if kk <= 32 else 1 << 21) and 9 <= kk <= 64 else ""),
				                     note="shard of %d MB of planes%s" % (n_s * 12 >> 20, " (fits the 126 MB L2)" if n_s * 12 < 120e6 else ""))}
			del ds
	# ---- config 4: 1024 images of 1920x1080, k=8 RGB k-means, images partitioned over the N GPUs (no collective) ----
	i0, i1 = partition_batch(1024, world, rank)
	nimg, n4, K8, iters = i1 - i0, 1920 * 1080, 8, 20
	g = torch.Generator(device=eng.dev)
	g.manual_seed(4000 + rank)
	batch = torch.randint(0, 256, (nimg, n4, 4), dtype=torch.uint8, device=eng.dev, generator=g)
	batch[:, :, 3] = 255
	c = [(batch[:, :K8, :3].double() + torch.arange(K8, device=eng.dev, dtype=torch.float64)[None, :, None] * 0.01).contiguous(),
	     torch.zeros((nimg, K8, 3), dtype=torch.float64, device=eng.dev)]
	lab4 = torch.empty((nimg, n4), dtype=torch.uint8, device=eng.dev)
	sums = torch.zeros((nimg, K8, 3), dtype=torch.float64, device=eng.dev)
	cnts = torch.zeros((nimg, K8), dtype=torch.float64, device=eng.dev)
	st4 = torch.zeros((nimg, 4), dtype=torch.float64, device=eng.dev)

	class _Batch:
		cur = 0

		def iterate(self):
			eng._call("cs_lloyd_iter_rgba8_batched", batch.data_ptr(), n4, nimg, -1, c[self.cur].data_ptr(), K8, lab4.data_ptr(),
			          sums.data_ptr(), cnts.data_ptr(), c[self.cur ^ 1].data_ptr(), st4.data_ptr(), _ffi.CS_LLOYD_EXACT_TIES)
			self.cur ^= 1

	bdrv = _Batch()
	for _ in range(3):
		bdrv.iterate()
	ms = timed_steps(bdrv, iters, lead=0) / iters
	ok4 = bool((cnts.sum(dim=1) == float(n4)).all().item())
	out["c4_batch_1024x1080p_k8"] = {
		"value": round(1024 * n4 / (ms * 1e-3) / 1e6, 1), "unit": UNIT + " per iteration, whole batch", "ms_per_step": round(ms, 5),
		"images_per_gpu": nimg, "label_mode": "exact_ties", "labels_written": True, "counts_sum_ok": ok4,
		"what": "ONE batched launch per Lloyd iteration over this GPU's share of the batch (gridDim.y = image); packed RGBA8 features",
		"roofline": roofline(ms, nimg * n4, 5.0, "lloyd_kernel<8, RGBA8> batched", note="4 B/px read + 1 B/px label written")}
	del batch, lab4, c, sums, cnts, st4
	torch.cuda.empty_cache()
	# ---- config 5: median-cut / octree / posterize, 256 colours, 16 MP (1 GPU; bit-exact integer path) ----
	if world == 1:
		rng = np.random.default_rng(5)
		img5 = np.dstack([rng.integers(0, 256, (4096, 4096, 3), dtype=np.uint8), np.full((4096, 4096), 255, np.uint8)])
		crop = np.ascontiguousarray(img5[:1024, :1024])
		c5 = {}
		for name, fn in (("median_cut", cs.simplify_colors_median_cut), ("octree", cs.simplify_colors_octree),
		                 ("threshold", cs.simplify_colors_threshold)):
			fn(img5, 256)
			torch.cuda.synchronize()
			t0 = time.perf_counter()
			fn(img5, 256)
			torch.cuda.synchronize()
			dt = time.perf_counter() - t0
			# bit-exactness against the library call the reference makes, on a 1 MP crop (Pillow on 16 MP takes ~9 s)
			o, p = fn(crop, 256)
			if name == "threshold":
				step = 256 // int(np.ceil(np.cbrt(256)))
				ro = crop.copy()
				ro[:, :, :3] = (crop[:, :, :3] // step) * step
				rp = np.unique(ro[:, :, :3].reshape(-1, 3), axis=0)[:256]
			else:
				from PIL import Image

				im = Image.fromarray(np.ascontiguousarray(crop[:, :, :3])).quantize(colors=256, method=Image.Quantize.MEDIANCUT)
				rp = np.array(im.getpalette()).reshape(-1, 3)[:256]
				ro = np.dstack([np.array(im.convert("RGB")), crop[:, :, 3]])
			c5[name] = {"ms_per_call_16mp_host_in_host_out": round(dt * 1e3, 2), "mpix_s": round(16.777216 / dt, 1),
			            "image_bit_exact": bool(np.array_equal(o, ro)), "palette_bit_exact": bool(np.array_equal(p, rp)),
			            "bit_exact_checked_on": "1024x1024 crop vs Pillow MEDIANCUT / NumPy posterize (the reference's calls)"}
		out["c5_integer_paths_16mp_k256"] = c5
	# ---- config 1: simplify_colors_kmeans(working_image_cleaned, 16) on the real 1024^2 input (1 GPU; latency) ----
	# ---- config 2: simplify_colors_perceptual_fast(3840x2160, 16), reference defaults, and its fit="full" variant ----
	if world == 1:
		def best_of(fn, reps=5):
			fn()
			best = None
			for _ in range(reps):
				torch.cuda.synchronize()
				t0 = time.perf_counter()
				r = fn()
				torch.cuda.synchronize()
				dt = time.perf_counter() - t0
				best = dt if best is None or dt < best else best
			return best, r

		try:
			gw = np.load(ROOT / "tests" / "golden" / "working_image_cleaned.npz")
			ge = np.load(ROOT / "tests" / "golden" / "reference_entry_points.npz")
			rgb1 = gw["colours"][gw["index"]]
			img1 = np.ascontiguousarray(np.dstack([rgb1, np.full(rgb1.shape[:2], 255, np.uint8)]))
			dt1, (o1, p1) = best_of(lambda: cs.simplify_colors_kmeans(img1, 16))
			ref1 = ge["bmp__kmeans_16__palette"]
			d1 = p1.astype(int) - ref1.astype(int) if p1.shape == ref1.shape else None
			dt1m, (_, p1m) = best_of(lambda: cs.simplify_colors_median_cut(img1, 16))
			out["c1_kmeans16_working_image_1024sq"] = {
				"ms_per_call_host_in_host_out": round(dt1 * 1e3, 3), "palette_rows": int(len(p1)),
				"palette_equal_or_plus1_vs_reference": bool(d1 is not None and ((d1 == 0) | (d1 == 1)).all()),
				"median_cut_16_ms": round(dt1m * 1e3, 3),
				"median_cut_palette_bit_exact_vs_reference": bool(np.array_equal(p1m, ge["bmp__median_cut_16__palette"])),
				"what": "the reference's own 1024x1024 nine-colour input (tests/golden/working_image_cleaned.npz), 10 k-means++ "
				        "initialisations + Lloyd loops on the device, upload and download inside; the palettes are compared with what the "
				        "unmodified reference returned (tests/golden/reference_entry_points.npz; +1: DESIGN.md section 4, deviation 2)"}
		except FileNotFoundError as exc:
			out["c1_kmeans16_working_image_1024sq"] = {"unavailable": str(exc)}
		img2 = np.dstack([np.random.default_rng(2).integers(0, 256, (2160, 3840, 3), dtype=np.uint8), np.full((2160, 3840), 255, np.uint8)])
		n2 = 2160 * 3840

		def c2_call(**kw):
			np.random.seed(2)  # the reference draws its colour sample from NumPy's global generator
			return cs.simplify_colors_perceptual_fast(img2, 16, **kw)

		dt2, (o2, p2) = best_of(c2_call, reps=3)
		dt2f, (o2f, p2f) = best_of(lambda: c2_call(fit="full", max_iter=20, tol=-1.0), reps=3)
		d2 = eng.upload_rgba(img2)
		pl2 = eng.rgba_to_lab(d2)
		lab2 = torch.empty(n2, dtype=torch.uint8, device=eng.dev)
		dr2 = make_gpu_lloyd(eng, pl2, n2, 16, labels=lab2, exact=True)
		dr2.set_centers(initial_centers(16))
		for _ in range(5):
			dr2.iterate()
		ms2 = timed_steps(dr2, 50) / 50
		out["c2_perceptual_fast_4k_k16"] = {
			"ms_per_call_host_in_host_out": round(dt2 * 1e3, 2), "mpix_s": round(n2 / dt2 / 1e6, 1), "palette_rows": int(len(p2)),
			"fit_full_20_iterations_ms_per_call": round(dt2f * 1e3, 2),
			"what": "simplify_colors_perceptual_fast(3840x2160 uniform-random image, 16) with the reference's defaults (palette from a "
			        "<= 5000-colour sample, then the per-pixel LAB remap on the device), and the same call with fit='full' (20 exact "
			        "Lloyd iterations over all 8.3 MP from the sample fit's centres); ordinary NumPy arrays in and out",
			"lloyd_iteration_8mp": {"ms_per_step": round(ms2, 5), "value": round(n2 / (ms2 * 1e-3) / 1e6, 1), "unit": UNIT,
			                        "roofline": roofline(ms2, n2, BYTES_PER_PX, "lloyd_kernel<16> (full walk, exact labels)",
			                                             note="100 MB of planes: fits the 126 MB L2")}}
		del dr2, pl2, d2, lab2
	return out


if __name__ == "__main__":
	main()
