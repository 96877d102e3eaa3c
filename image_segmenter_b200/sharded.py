"""Multi-GPU Lloyd iterations: one process per GPU, pixels sharded by contiguous row blocks.

The per-pixel work is independent; the only exchange is the (K*3 sums + K counts) fp64 partial of
each rank once per iteration.  Two exchange back-ends:
  * "nccl"  — `torch.distributed.all_reduce(SUM)` of the 4K-double buffer, then the finalize kernel
              (the required baseline; with the gloo backend this same code path runs on CPU in the
              world_size-2 tests, with the local step supplied by the test);
  * "p2p"   — cs_lloyd_iter_f32_mg (csrc/lloyd.cu + csrc/mg.cu): the last CTA of every rank's step kernel
              stores its partial into every peer's mailbox over NVLink (cudaIpc-mapped peer memory), waits
              for the peers' partials of the same epoch, sums them in rank order and runs the M-step tail —
              one kernel per iteration, no collective launch.
Every rank ends each iteration with bit-identical centres (same values summed in the same order),
so no broadcast is needed and convergence decisions agree without further communication.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np


def shard_rows(height: int, world: int, rank: int) -> tuple[int, int]:
	"""Contiguous row block [r0, r1) of rank `rank`; the first height % world ranks get one extra row."""
	base, extra = divmod(int(height), int(world))
	r0 = rank * base + min(rank, extra)
	return r0, r0 + base + (1 if rank < extra else 0)


@dataclass
class ShardedResult:
	centers: np.ndarray
	n_iter: int
	shift2: float


class ShardedLloyd:
	"""Loop control + exchange around a local step.

	local_step(c_in, acc)            fills acc[0:3K] with this rank's per-cluster sums and acc[3K:4K] with
	                                 its counts for the centres c_in (K x 3 fp64 tensor); in a relocation
	                                 redo it must also leave this rank's labels where local_farthest finds them;
	finalize(acc, c_in, c_out, st)   M-step tail on the all-reduced acc: c_out = new centres, st[0] = sum of
	                                 squared centre shifts, st[1] = number of empty clusters;
	local_farthest(c_in, prev)       (optional) this rank's next relocation candidate after the previous
	                                 global pick `prev` = (distance, global index) or None:
	                                 -> (distance, global index or -1, x, y, z, label)
	                                 (cs_lloyd_farthest_f32 in the product, the oracle in the gloo tests);
	check_error()                    (optional) raises when the exchange reported a failure.
	Tensors live wherever the caller allocates them (CUDA for the product, CPU for the gloo tests).

	Empty clusters (sklearn/cluster/_k_means_common.pyx:167-211): `run` reads st[1] at every check; it is
	bit-identical on all ranks, so every rank takes the same branch.  An iteration that left a cluster
	empty is REDONE with relocation — local step, all-reduce, then for every empty cluster in index order
	the globally farthest pixel (all-gather of the ranks' candidates; distance descending, global index
	ascending — the picks of the unsharded run) moves into it — before the M-step tail.  Without a
	`local_farthest` hook the run raises instead of continuing with sklearn's un-relocated average."""

	def __init__(self, K: int, local_step: Callable, finalize: Callable, *, device, group=None,
	             check_every: int = 1, local_farthest: Optional[Callable] = None,
	             check_error: Optional[Callable] = None):
		import torch
		import torch.distributed as dist

		self.K = int(K)
		self.local_step, self.finalize = local_step, finalize
		self.local_farthest, self.check_error = local_farthest, check_error
		self.group = group
		self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
		self.check_every = max(1, int(check_every))
		self.acc = torch.zeros(4 * self.K, dtype=torch.float64, device=device)
		self.c = [torch.zeros((self.K, 3), dtype=torch.float64, device=device) for _ in range(2)]
		self.stats = torch.zeros(4, dtype=torch.float64, device=device)
		self.cur = 0
		self.n_relocated = 0

	def set_centers(self, centers: np.ndarray):
		import torch

		self.c[self.cur].copy_(torch.from_numpy(np.ascontiguousarray(centers, dtype=np.float64)).reshape(self.K, 3))

	def _unfused_iterate(self, relocate: bool = False):
		import torch.distributed as dist

		c_in, c_out = self.c[self.cur], self.c[self.cur ^ 1]
		self.local_step(c_in, self.acc)
		if self.world > 1:
			dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=self.group)
		if relocate:
			self._relocate(c_in)
		self.finalize(self.acc, c_in, c_out, self.stats)
		self.cur ^= 1

	def iterate(self):
		"""One Lloyd iteration across all ranks (asynchronous on CUDA).  make_gpu_lloyd replaces this by
		the fused kernels; `_unfused_iterate` stays the relocation path."""
		self._unfused_iterate()

	def _relocate(self, c_in):
		"""Distributed _relocate_empty_clusters_dense on the all-reduced self.acc (identical on all ranks)."""
		import torch
		import torch.distributed as dist

		K = self.K
		acc = self.acc.cpu().numpy().copy()
		sums, counts = acc[:3 * K].reshape(K, 3), acc[3 * K:]
		empty = np.nonzero(counts == 0)[0]
		prev = None
		for e in empty:
			rec = np.asarray(self.local_farthest(c_in, prev), dtype=np.float64)
			if self.world > 1:
				t = torch.from_numpy(rec).to(self.acc.device)
				g = [torch.empty_like(t) for _ in range(self.world)]
				dist.all_gather(g, t, group=self.group)
				recs = np.stack([x.cpu().numpy() for x in g])
			else:
				recs = rec[None]
			recs = recs[recs[:, 1] >= 0]
			if len(recs) == 0:
				break
			# first in (distance descending, global index ascending)
			best = recs[np.lexsort((recs[:, 1], -recs[:, 0]))[0]]
			if prev is None and best[0] == 0.0:
				break  # np.max(distances) == 0: sklearn returns without relocating
			old = int(best[5])
			sums[old] -= best[2:5]
			sums[e] = best[2:5]
			counts[e] = 1.0
			counts[old] -= 1.0
			prev = (float(best[0]), int(best[1]))
			self.n_relocated += 1
		self.acc.copy_(torch.from_numpy(acc))

	def run(self, centers: np.ndarray, max_iter: int, tol: float = 0.0) -> ShardedResult:
		self.set_centers(centers)
		it, shift2 = 0, float("inf")
		ck_it, ck_c = 0, self.c[self.cur].clone()  # state at the last check (to rewind to)
		every = self.check_every
		while it < max_iter:
			self.iterate()
			it += 1
			if it % every == 0 or it == max_iter:
				st = self.stats.cpu().numpy()  # identical on every rank
				shift2 = float(st[0])
				if self.check_error is not None and not np.isfinite(shift2):
					self.check_error()
				if st[1] > 0:
					if self.local_farthest is None:
						raise RuntimeError(f"iteration {it}: {int(st[1])} empty cluster(s) and no relocation hook")
					# rewind to the last checked state and replay one iteration at a time: the first one
					# that leaves a cluster empty is redone with relocation
					self.c[self.cur].copy_(ck_c)
					it = ck_it
					while True:
						self.iterate()
						it += 1
						st = self.stats.cpu().numpy()
						if st[1] > 0:
							self.cur ^= 1  # back to the centres that iteration started from
							self._unfused_iterate(relocate=True)
							st = self.stats.cpu().numpy()
							break
					shift2 = float(st[0])
				ck_it, ck_c = it, self.c[self.cur].clone()
				if shift2 <= tol:
					break
		return ShardedResult(self.c[self.cur].cpu().numpy().copy(), it, shift2)


def connect_mailboxes(eng, group=None) -> None:
	"""Create this rank's mailbox, all-gather the 64-byte cudaIpc handles, map the peers' mailboxes.
	Idempotent per engine; collective (every rank must call it)."""
	import ctypes as C

	import torch
	import torch.distributed as dist

	from . import _ffi

	if getattr(eng, "_mg_connected", False):
		return
	world, rank = dist.get_world_size(group), dist.get_rank(group)
	buf = (C.c_ubyte * 64)()
	_ffi.check(eng.ctx.lib.cs_mg_create(eng.ctx.handle, world, rank, C.addressof(buf)), "cs_mg_create")
	mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=eng.dev)
	allh = torch.empty((world, 64), dtype=torch.uint8, device=eng.dev)
	dist.all_gather_into_tensor(allh, mine, group=group)
	handles = np.ascontiguousarray(allh.cpu().numpy())
	_ffi.check(eng.ctx.lib.cs_mg_connect(eng.ctx.handle, handles.ctypes.data), "cs_mg_connect")
	dist.barrier(group)  # nobody launches an exchanging kernel before every mailbox is mapped
	eng._mg_connected = True


def make_gpu_lloyd(eng, planes, n_local: int, K: int, *, labels=None, exact: bool = True, group=None,
                   x2max: Optional[float] = None, check_every: int = 1, exchange: str = "auto",
                   index_base: Optional[int] = None, box="auto", grid_policy: int = 0) -> ShardedLloyd:
	"""ShardedLloyd whose local step / finalize are the CUDA kernels: the fused cs_lloyd_iter_f32 with a
	single rank; with several ranks either cs_lloyd_step_f32 -> NCCL all_reduce -> cs_lloyd_finalize
	(exchange="nccl") or the single fused compute+exchange kernel cs_lloyd_iter_f32_mg (exchange="p2p",
	the default for world > 1).  `exact` (default) = CS_LLOYD_EXACT_TIES: labels equal the fp64 first
	minimum.  `index_base` = global index of this shard's first pixel (default: derived from the ranks'
	shard sizes when a relocation first needs it).

	Contract of the "p2p" exchange: every rank must issue the SAME number of exchanging launches over the
	life of its engine (each launch is one epoch of the mailbox protocol); `run` does, `iterate` callers
	must.  A rank that waits 20 s for a peer poisons its outputs with NaN; `run` then raises on every rank
	(cs_mg_error) instead of iterating on."""
	import torch.distributed as dist

	from . import _ffi

	flags = _ffi.CS_LLOYD_EXACT_TIES if exact else 0
	x2 = _ffi.CS_LAB_NORM2_MAX if x2max is None else float(x2max)
	if box == "auto":  # CIELAB planes from cs_rgba8_to_lab unless the caller says otherwise; see engine.KMeansGPU
		box = _ffi.CS_LAB_BOX if x2max is None else None
	lp = labels.data_ptr() if labels is not None else None
	p0, p1, p2 = planes[0].data_ptr(), planes[1].data_ptr(), planes[2].data_ptr()
	world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
	K = int(K)

	def local_step(c_in, acc):
		eng.set_feature_box(box, grid_policy)
		eng._call("cs_lloyd_step_f32", p0, p1, p2, n_local, c_in.data_ptr(), K, lp, acc.data_ptr(),
		          acc.data_ptr() + 3 * K * 8, None, x2, flags)

	def finalize(acc, c_in, c_out, stats):
		eng._call("cs_lloyd_finalize", acc.data_ptr(), acc.data_ptr() + 3 * K * 8, c_in.data_ptr(), K, c_out.data_ptr(),
		          stats.data_ptr())

	rank = dist.get_rank(group) if world > 1 else 0
	state = {"labels": labels, "base": int(index_base) if index_base is not None else None}

	def local_step_labels(c_in, acc):
		# relocation redo: the labels of THIS step are what cs_lloyd_farthest_f32 reads
		if state["labels"] is None:
			import torch

			state["labels"] = torch.empty((n_local + 3) & ~3, dtype=torch.uint8, device=eng.dev)
		eng.set_feature_box(box, grid_policy)
		eng._call("cs_lloyd_step_f32", p0, p1, p2, n_local, c_in.data_ptr(), K, state["labels"].data_ptr(), acc.data_ptr(),
		          acc.data_ptr() + 3 * K * 8, None, x2, _ffi.CS_LLOYD_EXACT_TIES)

	def local_farthest(c_in, prev):
		import ctypes as C

		if state["base"] is None:
			# global index of this shard's first pixel: exclusive prefix sum of the shard sizes
			import torch

			sizes = torch.zeros(max(world, 1), dtype=torch.int64, device=eng.dev)
			sizes[rank] = n_local
			if world > 1:
				dist.all_reduce(sizes, group=group)
			state["base"] = int(sizes[:rank].sum().item())
		none = (1 << 64) - 1
		prev2 = (C.c_uint64 * 2)(0, none)
		if prev is not None:
			prev2[0] = int(np.float64(prev[0]).view(np.uint64))
			prev2[1] = int(prev[1])
		out6 = (C.c_uint64 * 6)()
		eng._call("cs_lloyd_farthest_f32", p0, p1, p2, n_local, state["labels"].data_ptr(), c_in.data_ptr(), K,
		          state["base"], C.addressof(prev2), C.addressof(out6))
		u = np.array(list(out6), dtype=np.uint64)
		f = u.view(np.float64)
		return (float(f[0]), -1.0 if int(u[1]) == none else float(int(u[1])), float(f[2]), float(f[3]), float(f[4]),
		        float(int(u[5])))

	def check_error():
		if world > 1 and getattr(eng, "_mg_connected", False):
			import ctypes as C

			ep = C.c_ulonglong(0)
			_ffi.check(eng.ctx.lib.cs_mg_error(eng.ctx.handle, C.byref(ep)), "cs_mg_error")
			if ep.value:
				raise _ffi.ColorSimplifyError(
					f"multi-GPU exchange timed out at epoch {ep.value} on rank {rank}: a peer did not publish its partial "
					"(did every rank issue the same number of exchanging launches?)")
		raise _ffi.ColorSimplifyError("non-finite centre shift in the sharded Lloyd loop")

	drv = ShardedLloyd(K, local_step, finalize, device=eng.dev, group=group, check_every=check_every,
	                   local_farthest=local_farthest, check_error=check_error)
	_plain_unfused = drv._unfused_iterate

	def _unfused(relocate: bool = False):
		if relocate:  # the redo needs labels: swap in the labelled step for this one iteration
			drv.local_step = local_step_labels
			try:
				_plain_unfused(True)
			finally:
				drv.local_step = local_step
		else:
			_plain_unfused(False)

	drv._unfused_iterate = _unfused
	# the planes are never written while the driver lives, so every fused launch after the first may start
	# its prologue under the tail of whatever precedes it on the stream (CS_LLOYD_CHAINED)
	chain = [0]
	if exchange == "auto":
		exchange = "p2p" if world > 1 else "nccl"
	if world > 1 and exchange == "p2p":
		connect_mailboxes(eng, group)

		def fused_mg():
			c_in, c_out = drv.c[drv.cur], drv.c[drv.cur ^ 1]
			eng.set_feature_box(box, grid_policy)
			eng._call("cs_lloyd_iter_f32_mg", p0, p1, p2, n_local, c_in.data_ptr(), K, lp, drv.acc.data_ptr(),
			          drv.acc.data_ptr() + 3 * K * 8, c_out.data_ptr(), drv.stats.data_ptr(), x2, flags | chain[0])
			drv.cur ^= 1
			chain[0] = _ffi.CS_LLOYD_CHAINED

		drv.iterate = fused_mg
	if world == 1:
		def fused():
			c_in, c_out = drv.c[drv.cur], drv.c[drv.cur ^ 1]
			eng.set_feature_box(box, grid_policy)
			eng._call("cs_lloyd_iter_f32", p0, p1, p2, n_local, c_in.data_ptr(), K, lp, drv.acc.data_ptr(),
			          drv.acc.data_ptr() + 3 * K * 8, c_out.data_ptr(), drv.stats.data_ptr(), x2, flags | chain[0])
			drv.cur ^= 1
			chain[0] = _ffi.CS_LLOYD_CHAINED

		drv.iterate = fused
	return drv


# ---- integer paths across GPUs: row-sharded histogram quantisation ------------------
class ShardedMedianCut:
	"""Pillow MEDIANCUT on an image whose rows are sharded across ranks.

	local_hist()                     -> this rank's 2^24-bin colour histogram (int32 tensor)
	global steps (identical on every rank, from the all-reduced histogram): fold to <= 65 536 cells, host
	box tree, per-box sums, palette — then local_map(cell->box LUT, shift, palette) maps the local rows.
	The only exchange is ONE all_reduce(SUM) of the 64 MB histogram; every rank derives the same palette
	from the same summed histogram, so no broadcast follows.  The callables are the CUDA kernels in the
	product (`make_gpu_median_cut`) and oracle functions in the gloo CPU test."""

	def __init__(self, local_hist, palette_from_hist, local_map, *, group=None):
		import torch.distributed as dist

		self.local_hist, self.palette_from_hist, self.local_map = local_hist, palette_from_hist, local_map
		self.group = group
		self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

	def run(self, n_colors: int):
		import torch.distributed as dist

		hist = self.local_hist()
		if self.world > 1:
			dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)
		plan = self.palette_from_hist(hist, int(n_colors))
		return self.local_map(plan), plan


def make_gpu_median_cut(eng, d_rgba, preserve_alpha: bool = True, group=None) -> ShardedMedianCut:
	"""ShardedMedianCut over the CUDA kernels (cs_hist_rgb24 / fold / compact / box_sums / palette_map) and
	the host box tree, for this rank's rows `d_rgba` ((n,4) uint8 device tensor)."""
	import torch

	def local_hist():
		hist = torch.zeros(1 << 24, dtype=torch.int32, device=eng.dev)
		eng._call("cs_hist_rgb24", d_rgba.data_ptr(), d_rgba.shape[0], hist.data_ptr())
		return hist

	def palette_from_hist(hist, n_colors):
		return eng.median_cut_plan(hist, n_colors)

	def local_map(plan):
		return eng.median_cut_apply(d_rgba, plan, preserve_alpha)

	return ShardedMedianCut(local_hist, palette_from_hist, local_map, group=group)
