"""Host-side engine: device buffers (torch tensors as carriers) + thin wrappers over the C ABI.

Nothing here computes on the CPU: every per-pixel step is a call into libcolorsimplify.so.  The
wrappers keep argument order and meaning of include/colorsimplify.h; `KMeansGPU` is the Lloyd
driver that replaces sklearn's `_kmeans_single_lloyd` loop (sklearn/cluster/_kmeans.py:630-758)
and the best-of-n_init selection of `KMeans.fit` (:1506-1541).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _colorspace as cspace
from . import _ffi

EXACT = _ffi.CS_LLOYD_EXACT_TIES


def _torch():
	import torch

	return torch


def _sel(sel):
	"""(selection pixel tensor, mask_mode, min_bright) | None -> C arguments."""
	if sel is None:
		return None, 0, -1
	px, mm, mb = sel
	return px.data_ptr(), int(mm), int(mb)


class Engine:
	"""One per CUDA device.  Holds the cs_ctx and a few persistent small device tensors."""

	def __init__(self, device: int | None = None):
		torch = _torch()
		self.ctx = _ffi.get_context(device)
		self.dev = self.ctx.torch_device
		self.lut256 = torch.from_numpy(cspace.linear_lut256()).to(self.dev)
		self._bitmap24 = None
		self._bitmap32 = None

	# ---- plumbing ---------------------------------------------------------------------
	def _call(self, name, *args):
		self.ctx.call(name, *args, self.ctx.stream())

	def upload_rgba(self, rgba: np.ndarray):
		"""HxWx4 uint8 -> device (n,4) uint8 tensor (16-byte aligned by the allocator)."""
		torch = _torch()
		flat = np.ascontiguousarray(rgba).reshape(-1, 4)
		t = torch.from_numpy(flat)
		if flat.nbytes >= self.STAGED_UPLOAD_MIN_BYTES and not t.is_pinned():
			return self._staged_upload(flat)
		return t.to(self.dev, non_blocking=False)

	# ordinary (pageable) arrays of at least this size go up through the library's threaded staging ring
	STAGED_UPLOAD_MIN_BYTES = 8 << 20

	def _staged_upload(self, flat: np.ndarray):
		"""cs_host_upload: host threads copy the pageable array into page-locked buffers, 8 MB at a time, and each
		buffer leaves as one DMA — ~3x the rate of the driver's own staged pageable copy."""
		torch = _torch()
		d = torch.empty(flat.shape, dtype=torch.uint8, device=self.dev)
		self._call("cs_host_upload", flat.ctypes.data, flat.nbytes, d.data_ptr())
		return d

	def upload_rgba_lab(self, rgba: np.ndarray, chunks: int = 8):
		"""HxWx4 uint8 -> (device (n,4) uint8, fp32 LAB planes (3, n4)).  The image goes up in `chunks` pieces
		on a copy stream while the LAB conversion (K1) of the previous piece runs on the current stream
		(when the source is page-locked the copies are asynchronous DMAs and the two overlap)."""
		torch = _torch()
		flat_np = np.ascontiguousarray(rgba).reshape(-1, 4)
		flat = torch.from_numpy(flat_np)
		n = flat.shape[0]
		planes = torch.empty((3, (n + 3) & ~3), dtype=torch.float32, device=self.dev)
		if flat_np.nbytes >= self.STAGED_UPLOAD_MIN_BYTES and not flat.is_pinned():
			# pageable source: the staged upload (its DMAs are already pipelined with the host copies), then K1
			d = self._staged_upload(flat_np)
			self._call("cs_rgba8_to_lab", d.data_ptr(), n, self.lut256.data_ptr(), planes[0].data_ptr(),
			           planes[1].data_ptr(), planes[2].data_ptr())
			return d, planes
		d = torch.empty((n, 4), dtype=torch.uint8, device=self.dev)
		if not hasattr(self, "_copy_stream"):
			self._copy_stream = torch.cuda.Stream(device=self.dev)
		cur = torch.cuda.current_stream(self.dev)
		per = ((n + chunks - 1) // chunks + 3) & ~3
		self._copy_stream.wait_stream(cur)
		for o in range(0, n, per):
			m = min(per, n - o)
			with torch.cuda.stream(self._copy_stream):
				d[o:o + m].copy_(flat[o:o + m], non_blocking=True)
				ev = torch.cuda.Event()
				ev.record(self._copy_stream)
			cur.wait_event(ev)
			self._call("cs_rgba8_to_lab", d[o:o + m].data_ptr(), m, self.lut256.data_ptr(), planes[0, o:].data_ptr(),
			           planes[1, o:].data_ptr(), planes[2, o:].data_ptr())
		return d, planes

	def to_host(self, t) -> np.ndarray:
		"""Large device tensor -> NumPy array in page-locked memory (torch's caching host allocator): one DMA at
		PCIe rate instead of the driver's staged pageable copy."""
		torch = _torch()
		host = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
		host.copy_(t)
		return host.numpy()

	def pin(self, arr: np.ndarray) -> None:
		"""Page-lock a caller-owned C-contiguous array in place (cs_host_register): e.g. the NumPy view of a
		QImage's bits.  Uploads from / downloads into it then run at PCIe rate.  Pair with unpin()."""
		if not arr.flags["C_CONTIGUOUS"]:
			raise ValueError("only a C-contiguous array can be page-locked in place")
		_ffi.check(self.ctx.lib.cs_host_register(arr.ctypes.data, arr.nbytes), "cs_host_register")

	def unpin(self, arr: np.ndarray) -> None:
		_ffi.check(self.ctx.lib.cs_host_unregister(arr.ctypes.data), "cs_host_unregister")

	def set_feature_box(self, box, policy: int = 0) -> None:
		"""Vouch for an axis-aligned box ((lo0,lo1,lo2),(hi0,hi1,hi2)) around the planar-fp32 features of the
		next Lloyd launches (cs_lloyd_set_feature_box: enables the grid-filtered assignment), or clear it (None).
		policy 0: the library uses the grid path where it is faster (9 <= K <= 64); 1: wherever eligible."""
		_ffi.check(self.ctx.lib.cs_lloyd_set_grid_policy(self.ctx.handle, int(policy)), "cs_lloyd_set_grid_policy")
		if box is None:
			_ffi.check(self.ctx.lib.cs_lloyd_set_feature_box(self.ctx.handle, None, None), "cs_lloyd_set_feature_box")
			return
		lo, hi = (C.c_double * 3)(*map(float, box[0])), (C.c_double * 3)(*map(float, box[1]))
		_ffi.check(self.ctx.lib.cs_lloyd_set_feature_box(self.ctx.handle, C.addressof(lo), C.addressof(hi)),
		           "cs_lloyd_set_feature_box")

	def empty(self, shape, dtype):
		return _torch().empty(shape, dtype=dtype, device=self.dev)

	def zeros(self, shape, dtype):
		return _torch().zeros(shape, dtype=dtype, device=self.dev)

	def bitmap24(self):
		"""2^24-bit presence bitmap (2 MiB), zeroed."""
		torch = _torch()
		if self._bitmap24 is None:
			self._bitmap24 = torch.zeros(1 << 19, dtype=torch.int32, device=self.dev)
		else:
			self._bitmap24.zero_()
		return self._bitmap24

	def bitmap32(self):
		"""2^32-bit presence bitmap (512 MiB), zeroed."""
		torch = _torch()
		if self._bitmap32 is None:
			self._bitmap32 = torch.zeros(1 << 27, dtype=torch.int32, device=self.dev)
		else:
			self._bitmap32.zero_()
		return self._bitmap32

	def release_large_buffers(self):
		self._bitmap32 = None

	# ---- K1 / K9 ----------------------------------------------------------------------
	def rgba_to_lab(self, d_rgba):
		torch = _torch()
		n = d_rgba.shape[0]
		npad = (n + 3) & ~3
		planes = torch.empty((3, npad), dtype=torch.float32, device=self.dev)
		self._call("cs_rgba8_to_lab", d_rgba.data_ptr(), n, self.lut256.data_ptr(), planes[0].data_ptr(),
		           planes[1].data_ptr(), planes[2].data_ptr())
		return planes

	def rgba_to_lab_f64(self, d_rgba):
		out = _torch().empty((d_rgba.shape[0], 3), dtype=_torch().float64, device=self.dev)
		self._call("cs_rgba8_to_lab_f64", d_rgba.data_ptr(), d_rgba.shape[0], self.lut256.data_ptr(), out.data_ptr())
		return out

	def rgba_to_hsv(self, d_rgba):
		out = _torch().empty_like(d_rgba)
		self._call("cs_rgba8_to_hsv8", d_rgba.data_ptr(), d_rgba.shape[0], out.data_ptr())
		return out

	# ---- K8 ---------------------------------------------------------------------------
	def popcount(self, bitmap) -> int:
		out = self.zeros(1, _torch().int64)
		self._call("cs_bitmap_popcount", bitmap.data_ptr(), bitmap.numel(), out.data_ptr())
		return int(out.item())

	def mask_stats(self, d_px, min_bright: int, hsv: bool = False, want_unique: bool = False):
		"""-> (n_opaque, n_bright_hi, n_bright_lo, n_unique or None)."""
		acc = self.zeros(4, _torch().int64)
		bm = self.bitmap24() if want_unique else None
		self._call("cs_mask_stats_hsv8" if hsv else "cs_mask_stats_rgba8", d_px.data_ptr(), d_px.shape[0],
		           int(min_bright), bm.data_ptr() if bm is not None else None, acc.data_ptr())
		a = acc.cpu().numpy()
		return int(a[0]), int(a[1]), int(a[2]), (self.popcount(bm) if want_unique else None)

	def statistics(self, d_rgba):
		"""-> (n_unique_rgba, n_opaque, sums[3], sumsq[3]) as exact Python ints."""
		acc = self.zeros(8, _torch().int64)
		bm = self.bitmap32()
		self._call("cs_stats_rgba8", d_rgba.data_ptr(), d_rgba.shape[0], bm.data_ptr(), acc.data_ptr())
		a = [int(v) for v in acc.cpu().numpy()]
		return self.popcount(bm), a[0], a[1:4], a[4:7]

	# ---- K4 ---------------------------------------------------------------------------
	def set_remap_policy(self, policy: int = 0) -> None:
		"""cs_remap_set_policy: 0 = by image size (default), 1 = three-phase tiles, 2 = per-colour table,
		-1 = the direct kernel.  The labels are the same on every path."""
		_ffi.check(self.ctx.lib.cs_remap_set_policy(self.ctx.handle, int(policy)), "cs_remap_set_policy")

	def assign_remap(self, d_rgba, space: int, centers: np.ndarray, palette_u8: np.ndarray, preserve_alpha: bool,
	                 want_labels: bool = False):
		torch = _torch()
		K = int(centers.shape[0])
		d_c = torch.from_numpy(np.ascontiguousarray(centers, dtype=np.float64)).to(self.dev)
		d_p = torch.from_numpy(np.ascontiguousarray(palette_u8, dtype=np.uint8)).to(self.dev)
		out = torch.empty_like(d_rgba)
		lab = torch.empty(d_rgba.shape[0], dtype=torch.uint8, device=self.dev) if want_labels else None
		self._call("cs_assign_remap_rgba8", d_rgba.data_ptr(), d_rgba.shape[0], space, self.lut256.data_ptr(),
		           d_c.data_ptr(), d_p.data_ptr(), K, int(bool(preserve_alpha)), out.data_ptr(),
		           lab.data_ptr() if lab is not None else None)
		return out, lab

	def remap_labels(self, d_rgba, d_labels, palette_u8: np.ndarray, preserve_alpha: bool, sel=None):
		"""`sel` = (selection pixels, mask_mode, min_bright) of the masked step that made the labels."""
		torch = _torch()
		d_p = torch.from_numpy(np.ascontiguousarray(palette_u8, dtype=np.uint8)).to(self.dev)
		out = torch.empty_like(d_rgba)
		sp, mm, mb = _sel(sel)
		self._call("cs_remap_labels_rgba8", d_rgba.data_ptr(), d_labels.data_ptr(), d_rgba.shape[0], sp, mm, mb,
		           d_p.data_ptr(), int(palette_u8.shape[0]), int(bool(preserve_alpha)), out.data_ptr())
		return out

	def sum_by_label(self, d_rgba, d_labels, K: int, sel=None) -> np.ndarray:
		acc = self.zeros((K, 4), _torch().int64)
		sp, mm, mb = _sel(sel)
		self._call("cs_sum_by_label_rgba8", d_rgba.data_ptr(), d_labels.data_ptr(), d_rgba.shape[0], sp, mm, mb, K,
		           acc.data_ptr())
		return acc.cpu().numpy()

	def merge_labels(self, d_primary, d_fallback, n: int, sel=None):
		out = _torch().empty_like(d_primary)
		sp, mm, mb = _sel(sel)
		self._call("cs_merge_labels_u8", d_primary.data_ptr(), d_fallback.data_ptr(), int(n), sp, mm, mb, out.data_ptr())
		return out

	def same_clustering(self, l1, l2, n: int, sel=None) -> bool:
		"""_is_same_clustering (sklearn/cluster/_k_means_common.pyx:314-328): every label of l1 maps to
		a single label of l2 — decided from the 256 x 256 co-occurrence matrix the kernel marks."""
		mat = self.empty((256, 256), _torch().int32)
		sp, mm, mb = _sel(sel)
		self._call("cs_label_cooccurrence_u8", l1.data_ptr(), l2.data_ptr(), int(n), sp, mm, mb, mat.data_ptr())
		m = mat.cpu().numpy()
		return bool((m.sum(axis=1) <= 1).all())

	def nn_argmin_rows64(self, d_query, d_ref) -> np.ndarray:
		"""pairwise_distances_argmin_min indices for fp64 rows on the device (cs_nn_argmin_rows64)."""
		torch = _torch()
		q, r = d_query.contiguous(), d_ref.contiguous()
		out = torch.empty(q.shape[0], dtype=torch.int64, device=self.dev)
		self._call("cs_nn_argmin_rows64", q.data_ptr(), q.shape[0], r.data_ptr(), r.shape[0], out.data_ptr())
		return out.cpu().numpy()

	# ---- selection / sampling ---------------------------------------------------------
	def select_count(self, d_px, mask_mode: int, min_bright: int) -> int:
		cnt = self.zeros(1, _torch().int64)
		self._call("cs_select_compact_px8", d_px.data_ptr(), d_px.shape[0], int(mask_mode), int(min_bright), None, None, 0,
		           cnt.data_ptr())
		return int(cnt.item())

	def select_compact(self, d_px, mask_mode: int, min_bright: int, want_index: bool = False):
		"""-> (compacted (m,4) uint8 pixels, optional (m,) int64 source positions), order-preserving."""
		torch = _torch()
		m = self.select_count(d_px, mask_mode, min_bright)
		out = torch.empty((max(m, 1), 4), dtype=torch.uint8, device=self.dev)
		idx = torch.empty(max(m, 1), dtype=torch.int64, device=self.dev) if want_index else None
		cnt = self.zeros(1, torch.int64)
		self._call("cs_select_compact_px8", d_px.data_ptr(), d_px.shape[0], int(mask_mode), int(min_bright), out.data_ptr(),
		           idx.data_ptr() if idx is not None else None, m, cnt.data_ptr())
		return out[:m], (idx[:m] if idx is not None else None)

	def channel_hist(self, d_px, mask_mode: int, min_bright: int) -> np.ndarray:
		"""(3, 256) exact per-byte counts of the selected pixels."""
		h = self.zeros(768, _torch().int64)
		self._call("cs_channel_hist_px8", d_px.data_ptr(), d_px.shape[0], int(mask_mode), int(min_bright), h.data_ptr())
		return h.cpu().numpy().reshape(3, 256)

	def gather(self, d_px, index: np.ndarray):
		torch = _torch()
		d_i = torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(self.dev)
		out = torch.empty((len(index), 4), dtype=torch.uint8, device=self.dev)
		self._call("cs_gather_px8", d_px.data_ptr(), d_px.shape[0], d_i.data_ptr(), len(index), out.data_ptr())
		return out

	# ---- k-means++ seeding --------------------------------------------------------------
	def kmeanspp_seeds(self, cpx, lut64, K: int, n_init: int, seed: int = 42, rows=None):
		"""The n_init k-means++ initialisations KMeans(n_clusters=K, random_state=seed).fit would draw for
		the rows `cpx` (compacted selected pixels, (n,4) uint8 on the device; features = lut64[c][byte c]).

		Host: the RandomState stream and the per-round decisions of sklearn's _kmeans_plusplus
		(sklearn/cluster/_kmeans.py:180-278), drawn back to back for the n_init runs as KMeans.fit does
		(Lloyd consumes no randomness, :1506-1514).  Device: every O(N) step (cs_kpp_*).
		Returns (indices list of (K,) int64 arrays, centres list of (K,3) float64 arrays).
		`rows` (an (n,3) float64 device tensor; cpx / lut64 None): the samples are fp64 feature rows instead of
		packed pixels — the standardised CIELAB rows of simplify_colors_adaptive_distance.

		Documented deviations (probability ~N * 1e-13 per draw): distances are the direct fp64 formula on the
		uncentred features (sklearn: |x|^2 - 2 x.c + |c|^2 on mean-centred data), cumulative sums are formed
		per 4096-pixel tile, and for n > 2^22 the first index is floor(u * n) instead of numpy's
		searchsorted over the cumulative sum of n copies of 1/n."""
		torch = _torch()
		from sklearn.utils import check_random_state

		n = int(rows.shape[0] if rows is not None else cpx.shape[0])
		rs = check_random_state(seed)
		T = 2 + int(np.log(K))
		if rows is None:
			lut64 = np.ascontiguousarray(lut64, dtype=np.float64).reshape(3, 256)
			d_lut = torch.from_numpy(lut64.reshape(-1)).to(self.dev)
		p_px = cpx.data_ptr() if rows is None else None
		p_lut = d_lut.data_ptr() if rows is None else None
		p_rows = rows.data_ptr() if rows is not None else None

		def rows_at(index):
			return rows[torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(self.dev)].cpu().numpy()
		ntiles = (n + 4095) // 4096
		cdf = None
		if n <= (1 << 22):
			p = np.ones(n, dtype=np.float64) / np.float64(n)  # sample_weight / sample_weight.sum()
			cdf = p.cumsum()
			cdf /= cdf[-1]

		def feats(px_rows):
			return np.stack([lut64[c][px_rows[:, c]] for c in range(3)], axis=1)

		# The random numbers a fit consumes are fixed in count (Lloyd draws none): one random_sample() for the first
		# centre and T uniforms per later centre, init after init.  Drawing them up front lets the n_init
		# initialisations advance in LOCKSTEP, and — the uniforms of every round being on the device — lets a whole
		# round (draw, candidate potentials, pick, update) run without the host: three read-backs per group of
		# initialisations (first pixel, first tile sums, the finished seeds) instead of two per round.
		draws = [(rs.random_sample(), [rs.uniform(size=T) for _ in range(1, K)]) for _ in range(n_init)]
		G = n_init if n_init * n * 8 <= (1 << 30) else 1  # closest-distance arrays of a group: <= 1 GiB
		nblk_cap = 4 * self.ctx.sm_count + 8
		all_idx, all_cent = [], []
		nb = C.c_int(0)
		for g0 in range(0, n_init, G):
			grp = list(range(g0, min(n_init, g0 + G)))
			m = len(grp)
			closest = torch.empty((m, n), dtype=torch.float64, device=self.dev)
			tile_sums = torch.empty((m, ntiles), dtype=torch.float64, device=self.dev)
			block_pots = torch.empty((m, nblk_cap, 8), dtype=torch.float64, device=self.dev)
			d_pick = torch.empty(m, dtype=torch.int32, device=self.dev)
			d_cidx = torch.empty((m, 8), dtype=torch.int64, device=self.dev)
			d_idx_out = torch.empty((m, K), dtype=torch.int64, device=self.dev)  # slot 0 is filled on the host below
			d_cent_out = torch.empty((m, K, 3), dtype=torch.float64, device=self.dev)
			# first centre: random_state.choice(n_samples, p=sample_weight / sample_weight.sum())
			cids = []
			for j in grp:
				u = draws[j][0]
				cid = int(cdf.searchsorted(u, side="right")) if cdf is not None else min(int(u * n), n - 1)
				cids.append(min(cid, n - 1))
			cands_h = np.zeros((m, 8, 3), dtype=np.float64)  # candidate features of the round, per initialisation
			cands_h[:, 0] = rows_at(np.array(cids)) if rows is not None else feats(self.gather(cpx, np.array(cids)).cpu().numpy())
			d_cands = torch.from_numpy(cands_h).to(self.dev)
			self._call("cs_kpp_update_batched", p_px, n, p_lut, p_rows, d_cands.data_ptr(), None, 1,
			           closest.data_ptr(), tile_sums.data_ptr(), m)
			if K > 1:
				# pot = closest_dist_sq.sum() as NumPy forms it (pairwise over the tile sums), then the device keeps it
				d_pot = torch.from_numpy(np.array([float(t.sum()) for t in tile_sums.cpu().numpy()], dtype=np.float64)).to(self.dev)
				uni = np.zeros((K - 1, m, 8), dtype=np.float64)
				for i, j in enumerate(grp):
					for c in range(1, K):
						uni[c - 1, i, :T] = draws[j][1][c - 1]
				d_uni = torch.from_numpy(uni).to(self.dev)
				for c in range(1, K):
					self._call("cs_kpp_draw_batched", closest.data_ptr(), n, tile_sums.data_ptr(), d_pot.data_ptr(),
					           d_uni[c - 1].data_ptr(), T, p_px, p_lut, p_rows, d_cidx.data_ptr(), d_cands.data_ptr(), m)
					_ffi.check(self.ctx.lib.cs_kpp_eval_batched(self.ctx.handle, p_px, n, p_lut, p_rows, d_cands.data_ptr(), T,
					                                            closest.data_ptr(), block_pots.data_ptr(), nblk_cap, m, C.byref(nb),
					                                            self.ctx.stream()), "cs_kpp_eval_batched")
					# potentials, their first minimum, the winner's record and the update with it: all queued
					self._call("cs_kpp_pick_batched", block_pots.data_ptr(), nblk_cap, nb.value, T, m, d_pick.data_ptr(), d_pot.data_ptr())
					self._call("cs_kpp_record_batched", d_cidx.data_ptr(), d_cands.data_ptr(), d_pick.data_ptr(), c, K, m,
					           d_idx_out.data_ptr(), d_cent_out.data_ptr())
					self._call("cs_kpp_update_batched", p_px, n, p_lut, p_rows, d_cands.data_ptr(), d_pick.data_ptr(), 0,
					           closest.data_ptr(), tile_sums.data_ptr(), m)
			h_idx, h_cent = d_idx_out.cpu().numpy(), d_cent_out.cpu().numpy()
			for i in range(m):
				h_idx[i, 0] = cids[i]
				h_cent[i, 0] = cands_h[i, 0]
				all_idx.append(h_idx[i].copy())
				all_cent.append(h_cent[i].copy())
		return all_idx, all_cent

	# ---- K7 ---------------------------------------------------------------------------
	def posterize(self, d_rgba, step: int, preserve_alpha: bool):
		"""-> (device rgba out, sorted unique quantised colours (U,3) uint8)."""
		torch = _torch()
		out = torch.empty_like(d_rgba)
		bm = self.bitmap24()
		self._call("cs_posterize_rgba8", d_rgba.data_ptr(), d_rgba.shape[0], int(step), int(bool(preserve_alpha)),
		           out.data_ptr(), bm.data_ptr())
		# the only colours the output can hold are the multiples of `step` per channel: test exactly those bits of
		# the presence bitmap on the device and read back one byte per candidate (not the 2 MiB bitmap)
		lv = np.arange(0, 256, int(step), dtype=np.int64)
		cand = ((lv[:, None, None] << 16) | (lv[None, :, None] << 8) | lv[None, None, :]).reshape(-1)  # ascending == np.unique row order
		d_k = torch.from_numpy(cand).to(self.dev)
		hit = ((bm[d_k >> 5] >> (d_k & 31)) & 1).to(torch.uint8).cpu().numpy().astype(bool)
		keys = cand[hit]
		pal = np.stack([(keys >> 16) & 0xFF, (keys >> 8) & 0xFF, keys & 0xFF], axis=1).astype(np.uint8) if len(keys) else np.zeros((0, 3), np.uint8)
		return out, pal

	# ---- K5 / K6: median cut ------------------------------------------------------------
	def median_cut(self, d_rgba, n_colors: int, preserve_alpha: bool):
		"""Pillow MEDIANCUT on the RGB of every pixel -> (device rgba out, palette (P,3) uint8, device indices)."""
		torch = _torch()
		hist = torch.zeros(1 << 24, dtype=torch.int32, device=self.dev)
		self._call("cs_hist_rgb24", d_rgba.data_ptr(), d_rgba.shape[0], hist.data_ptr())
		plan = self.median_cut_plan(hist, n_colors)
		out, idx = self.median_cut_apply(d_rgba, plan, preserve_alpha)
		return out, plan["palette"], idx

	def median_cut_plan(self, hist, n_colors: int) -> dict:
		"""From a (possibly all-reduced) 2^24-bin histogram: cell fold at the smallest shift with <= 65536
		cells (create_pixel_hash), host box tree, per-box sums -> {shift, lut (device cell->box), palette}."""
		torch = _torch()
		# occupied-cell counts at shift 0 (the histogram itself) and at the folds 1..3, all queued before ONE
		# read-back.  Shift 3 always qualifies (2^15 cells <= 65 536), so deeper folds are never needed — and they
		# are the expensive ones (2^24 atomic adds into a handful of cells).
		ncells = torch.zeros(4, dtype=torch.int32, device=self.dev)
		ncells[0] = torch.count_nonzero(hist)
		folded = [hist]
		for sh in range(1, 4):
			c = torch.empty(1 << (3 * (8 - sh)), dtype=torch.int32, device=self.dev)
			self._call("cs_hist_fold", hist.data_ptr(), sh, c.data_ptr(), ncells[sh:].data_ptr())
			folded.append(c)
		counts_h = ncells.cpu().numpy()
		shift = int(np.argmax(counts_h <= 65536))  # smallest shift with <= 65 536 cells (create_pixel_hash)
		nc, cells = int(counts_h[shift]), folded[shift]
		del folded
		ncell = ncells[:1]
		keys = torch.empty(nc, dtype=torch.int32, device=self.dev)
		counts = torch.empty(nc, dtype=torch.int32, device=self.dev)
		self._call("cs_hist_compact", cells.data_ptr(), cells.numel(), keys.data_ptr(), counts.data_ptr(), nc,
		           ncell.data_ptr())
		h_keys = keys.cpu().numpy().view(np.uint32)
		h_counts = counts.cpu().numpy().view(np.uint32)
		h_box = np.zeros(nc, dtype=np.uint16)
		n_boxes = C.c_int(0)
		_ffi.check(self.ctx.lib.cs_median_cut_boxes(h_keys.ctypes.data, h_counts.ctypes.data, nc, shift, int(n_colors),
		                                           h_box.ctypes.data, C.byref(n_boxes)), "cs_median_cut_boxes")
		P = n_boxes.value
		lut = np.full(1 << (3 * (8 - shift)), 0xFFFF, dtype=np.uint16)
		lut[h_keys] = h_box
		d_lut = torch.from_numpy(lut.view(np.int16)).to(self.dev)
		acc = torch.zeros((P, 4), dtype=torch.int64, device=self.dev)
		self._call("cs_box_sums", hist.data_ptr(), d_lut.data_ptr(), shift, P, acc.data_ptr())
		a = acc.cpu().numpy().astype(np.uint64)
		# compute_palette_from_median_cut keeps sums and counts in uint32 and rounds (int)(0.5 + sum/count)
		s32 = (a[:, :3] & np.uint64(0xFFFFFFFF)).astype(np.float64)
		c32 = (a[:, 3] & np.uint64(0xFFFFFFFF)).astype(np.float64)
		pal = (0.5 + s32 / c32[:, None]).astype(np.int64).astype(np.uint8)
		return {"shift": shift, "lut": d_lut, "palette": pal}

	def median_cut_apply(self, d_rgba, plan: dict, preserve_alpha: bool):
		"""Nearest-palette map (Pillow's tie rule) of the pixels with a plan from median_cut_plan."""
		torch = _torch()
		n = d_rgba.shape[0]
		pal = plan["palette"]
		d_pal = torch.from_numpy(pal).to(self.dev)
		out = torch.empty_like(d_rgba)
		idx = torch.empty(n, dtype=torch.uint8, device=self.dev)
		self._call("cs_palette_map_rgba8", d_rgba.data_ptr(), n, plan["lut"].data_ptr(), plan["shift"], d_pal.data_ptr(),
		           int(len(pal)), int(bool(preserve_alpha)), out.data_ptr(), idx.data_ptr())
		return out, idx


@dataclass
class FitResult:
	labels: object  # device uint8 tensor (255 = masked pixel)
	centers: np.ndarray  # (K,3) float64
	inertia: float
	n_iter: int
	sums: np.ndarray = None  # (K,3) float64 per-cluster feature sums of the last M-step (centers = sums / counts)
	counts: np.ndarray = None  # (K,) float64 cluster sizes of the last M-step


class KMeansGPU:
	"""Lloyd iterations on the GPU from given initial centres.

	`kind` selects the feature source: "f32" (three planar fp32 tensors), "rgba8" (packed RGBA, RGB
	features, brightness mask) or "px8lut" (packed 4 x u8 pixels through 3 x 256 feature tables).
	"""

	def __init__(self, eng: Engine, kind: str, n: int, *, planes=None, px=None, lut3=None, mask_mode=0,
	             min_bright=-1, x2max=_ffi.CS_LAB_NORM2_MAX, exact: bool = True, box="auto", grid_policy: int = 0):
		torch = _torch()
		self.eng, self.kind, self.n = eng, kind, int(n)
		# feature box for the grid-filtered assignment: CIELAB planes from cs_rgba8_to_lab by default.  The library
		# takes that path only where it beats the full walk (9 <= K <= 64; grid_policy=1 forces it from K = 4,
		# profiles/r2_grid_assignment.md)
		self.box = (_ffi.CS_LAB_BOX if (kind == "f32" and x2max == _ffi.CS_LAB_NORM2_MAX) else None) if box == "auto" else box
		self.grid_policy = int(grid_policy)
		self.planes, self.px, self.lut3 = planes, px, lut3
		self.mask_mode, self.min_bright, self.x2max = int(mask_mode), int(min_bright), float(x2max)
		self.flags = EXACT if exact else 0
		self.labels = torch.empty((self.n + 3) & ~3, dtype=torch.uint8, device=eng.dev)
		# the pixels whose selection defines which labels are real (None for unmasked planes)
		self.sel = (px, self.mask_mode, self.min_bright) if px is not None else None

	def _step(self, d_cin, K, d_sums, d_counts, labels=None, inertia=None, d_cout=None, d_stats=None, flags=None):
		e, n = self.eng, self.n
		e.set_feature_box(self.box, self.grid_policy)
		fl = self.flags if flags is None else flags
		lp = labels.data_ptr() if labels is not None else None
		ip = inertia.data_ptr() if inertia is not None else None
		if self.kind == "f32":
			p = self.planes
			if d_cout is not None:
				e._call("cs_lloyd_iter_f32", p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), n, d_cin.data_ptr(), K, lp,
				        d_sums.data_ptr(), d_counts.data_ptr(), d_cout.data_ptr(), d_stats.data_ptr(), self.x2max, fl)
			else:
				e._call("cs_lloyd_step_f32", p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), n, d_cin.data_ptr(), K, lp,
				        d_sums.data_ptr(), d_counts.data_ptr(), ip, self.x2max, fl)
		else:
			lut = self.lut3.data_ptr() if self.lut3 is not None else None
			if lut is None:
				if d_cout is not None:
					e._call("cs_lloyd_iter_rgba8", self.px.data_ptr(), n, self.min_bright, d_cin.data_ptr(), K, lp,
					        d_sums.data_ptr(), d_counts.data_ptr(), d_cout.data_ptr(), d_stats.data_ptr(), fl)
				else:
					e._call("cs_lloyd_step_rgba8", self.px.data_ptr(), n, self.min_bright, d_cin.data_ptr(), K, lp,
					        d_sums.data_ptr(), d_counts.data_ptr(), ip, fl)
			else:
				e._call("cs_lloyd_step_px8lut", self.px.data_ptr(), n, lut, self.mask_mode, self.min_bright, self.x2max,
				        d_cin.data_ptr(), K, lp, d_sums.data_ptr(), d_counts.data_ptr(), ip,
				        d_cout.data_ptr() if d_cout is not None else None,
				        d_stats.data_ptr() if d_stats is not None else None, fl)

	def _relocate(self, d_cold, K, d_sums, d_counts):
		e, n = self.eng, self.n
		if self.kind == "f32":
			p = self.planes
			e._call("cs_lloyd_relocate_f32", p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), n, self.labels.data_ptr(),
			        d_cold.data_ptr(), K, d_sums.data_ptr(), d_counts.data_ptr())
		else:
			e._call("cs_lloyd_relocate_px8", self.px.data_ptr(), n, self.lut3.data_ptr() if self.lut3 is not None else None,
			        self.labels.data_ptr(), d_cold.data_ptr(), K, d_sums.data_ptr(), d_counts.data_ptr())

	def _run(self, d_a, d_b, K, d_sums, d_counts, d_stats, n_launch, d_ctl):
		"""Queue n_launch fused iterations (a -> b -> a ...) with device-side loop control."""
		e, n = self.eng, self.n
		e.set_feature_box(self.box, self.grid_policy)
		if self.kind == "f32":
			p = self.planes
			e._call("cs_lloyd_run_f32", p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), n, d_a.data_ptr(), d_b.data_ptr(), K,
			        d_sums.data_ptr(), d_counts.data_ptr(), d_stats.data_ptr(), self.x2max, self.flags, int(n_launch),
			        d_ctl.data_ptr())
		else:
			lut = self.lut3.data_ptr() if self.lut3 is not None else None
			e._call("cs_lloyd_run_px8", self.px.data_ptr(), n, lut, self.mask_mode, self.min_bright, self.x2max, d_a.data_ptr(),
			        d_b.data_ptr(), K, d_sums.data_ptr(), d_counts.data_ptr(), d_stats.data_ptr(), self.flags, int(n_launch),
			        d_ctl.data_ptr())

	def _loop_many(self, inits, max_iter: int, tol: float, batch: int):
		"""The Lloyd loop of _kmeans_single_lloyd (sklearn/cluster/_kmeans.py:705-738) for SEVERAL initialisations
		in lockstep: iterations are queued in batches with the convergence / empty-cluster test on the device
		(cs_lloyd_run_*), one batch per initialisation, then ONE read-back of all control blocks — a host round
		trip per batch, not per batch and initialisation.  -> list of (device centres, iterations done, device
		sums, device counts)."""
		torch = _torch()
		e = self.eng
		m = len(inits)
		K = int(inits[0].shape[0])
		c0 = torch.from_numpy(np.ascontiguousarray(np.stack(inits), dtype=np.float64)).to(e.dev)  # ONE upload for all of them
		c1 = e.zeros((m, K, 3), torch.float64)
		c = [[c0[i], c1[i]] for i in range(m)]
		sums, counts = e.zeros((m, K, 3), torch.float64), e.zeros((m, K), torch.float64)
		stats = e.zeros((m, 4), torch.float64)
		ctl = torch.tensor([[0.0, 0.0, float(tol), 0.0]] * m, dtype=torch.float64, device=e.dev)
		cur, it, live = [0] * m, [0] * m, [max_iter > 0] * m
		while any(live):
			for i in range(m):
				if live[i]:
					self._run(c[i][cur[i]], c[i][cur[i] ^ 1], K, sums[i], counts[i], stats[i], min(int(batch), max_iter - it[i]), ctl[i])
			h = ctl.cpu().numpy()
			for i in range(m):
				if not live[i]:
					continue
				done = int(h[i, 1]) - it[i]
				cur[i] ^= done & 1
				it[i] += done
				if h[i, 0] == 1.0:
					live[i] = False
				elif h[i, 0] == 2.0:  # iteration it+1 found an empty cluster: redo it with labels, relocate, finish the M-step
					ci = c[i]
					self._step(ci[cur[i]], K, sums[i], counts[i], labels=self.labels)
					self._relocate(ci[cur[i]], K, sums[i], counts[i])
					e._call("cs_lloyd_finalize", sums[i].data_ptr(), counts[i].data_ptr(), ci[cur[i]].data_ptr(), K,
					        ci[cur[i] ^ 1].data_ptr(), stats[i].data_ptr())
					st = stats[i].cpu().numpy()
					cur[i] ^= 1
					it[i] += 1
					ctl[i].copy_(torch.tensor([0.0, float(it[i]), float(tol), 0.0], dtype=torch.float64))
					if st[0] <= tol:
						live[i] = False
				if it[i] >= max_iter:
					live[i] = False
		return [(c[i][cur[i]], it[i], sums[i], counts[i]) for i in range(m)]

	def _loop(self, init: np.ndarray, max_iter: int, tol: float, batch: int):
		return self._loop_many([init], max_iter, tol, batch)[0]

	def fit_centers(self, init: np.ndarray, max_iter: int = 300, tol: float = 0.0, batch: int = 10) -> np.ndarray:
		"""The loop alone: final centres (K,3) float64, no E-step (the caller assigns with its own kernel)."""
		return self._loop(init, max_iter, tol, batch)[0].cpu().numpy()

	def fit_single(self, init: np.ndarray, max_iter: int = 300, tol: float = 0.0, batch: int = 8) -> FitResult:
		"""_kmeans_single_lloyd: iterate until sum shift^2 <= tol (labels unchanged implies shift 0)
		or max_iter, then one E-step on the final centres for labels and inertia."""
		torch = _torch()
		e = self.eng
		K = int(init.shape[0])
		c_fin, it, sums, counts = self._loop(init, max_iter, tol, batch)
		inert = e.zeros(1, torch.float64)
		# final E-step on the final centres (labels + inertia); its sums go to scratch so that `sums` / `counts`
		# stay those of the M-step that PRODUCED the final centres (they differ after a tol stop)
		s2, c2 = torch.empty_like(sums), torch.empty_like(counts)
		self._step(c_fin, K, s2, c2, labels=self.labels, inertia=inert)
		return FitResult(self.labels, c_fin.cpu().numpy(), float(inert.item()), it, sums.cpu().numpy(), counts.cpu().numpy())

	def fit_best(self, inits, max_iter: int = 300, tol: float = 0.0, batch: int = 8) -> FitResult:
		"""Best of several initialisations by inertia, as KMeans.fit (sklearn/cluster/_kmeans.py:1506-1541).
		The runs advance in lockstep (`_loop_many`), their final E-steps are queued together and the inertias
		come back in one read; the selection then walks the runs in order exactly as sklearn does."""
		torch = _torch()
		e = self.eng
		inits = list(inits)
		m, K = len(inits), int(inits[0].shape[0])
		runs = self._loop_many(inits, max_iter, tol, batch)
		labs = [self.labels] + [torch.empty_like(self.labels) for _ in range(m - 1)]
		inert = e.zeros(m, torch.float64)
		s2, c2 = e.zeros((K, 3), torch.float64), e.zeros(K, torch.float64)
		for i, (c_fin, _, _, _) in enumerate(runs):
			# final E-step on the final centres (labels + inertia); its sums go to scratch so that `sums` / `counts`
			# stay those of the M-step that PRODUCED the final centres (they differ after a tol stop)
			self._step(c_fin, K, s2, c2, labels=labs[i], inertia=inert[i:i + 1])
		h_in = inert.cpu().numpy()
		best = None
		for i in range(m):
			if best is None or (h_in[i] < h_in[best] and not self.eng.same_clustering(labs[i], labs[best], self.n, self.sel)):
				best = i
		c_fin, it, sums, counts = runs[best]
		return FitResult(labs[best], c_fin.cpu().numpy(), float(h_in[best]), it, sums.cpu().numpy(), counts.cpu().numpy())


class KMeansRows64:
	"""KMeans(n_clusters=K, random_state=42, n_init=10).fit_predict(X) for n x 3 fp64 rows on the device: the
	full-N fallback fit of simplify_colors_adaptive_distance (reference color_simplify.py:809-814).  k-means++
	seeding (Engine.kmeanspp_seeds on rows), then per run the loop of _kmeans_single_lloyd
	(sklearn/cluster/_kmeans.py:705-738): fp64 first-minimum labels (cs_lloyd_step_rows64), per-cluster sums in a
	fixed order (cs_sum_by_label_rows64), M-step tail (cs_lloyd_finalize); stop when the labels repeat (strict)
	or sum(shift^2) <= tol = 1e-4 * mean variance; best run by inertia.  Small-image code (DBSCAN limits the
	caller to ~10^5 rows): one host round trip per iteration.  Empty clusters are relocated on the host from
	the device labels (rare)."""

	def __init__(self, eng: Engine, rows):
		self.eng, self.rows, self.n = eng, rows, int(rows.shape[0])

	def _run(self, init: np.ndarray, max_iter: int, tol: float):
		torch = _torch()
		e, n = self.eng, self.n
		K = int(init.shape[0])
		c = [torch.from_numpy(np.ascontiguousarray(init, dtype=np.float64)).to(e.dev), e.zeros((K, 3), torch.float64)]
		lab = [e.empty(n, torch.int32), e.empty(n, torch.int32)]
		sums, counts = e.zeros((K, 3), torch.float64), e.zeros(K, torch.float64)
		stats, inert = e.zeros(4, torch.float64), e.zeros(1, torch.float64)
		cur = 0
		lab[1].fill_(-1)
		for it in range(1, max_iter + 1):
			new, old = lab[it & 1 ^ 1], lab[it & 1]
			e._call("cs_lloyd_step_rows64", self.rows.data_ptr(), n, c[cur].data_ptr(), K, new.data_ptr(), inert.data_ptr())
			e._call("cs_sum_by_label_rows64", self.rows.data_ptr(), n, new.data_ptr(), K, sums.data_ptr(), counts.data_ptr())
			if bool((counts == 0).any().item()):
				self._relocate_host(new, c[cur], sums, counts)
			e._call("cs_lloyd_finalize", sums.data_ptr(), counts.data_ptr(), c[cur].data_ptr(), K, c[cur ^ 1].data_ptr(), stats.data_ptr())
			cur ^= 1
			if bool(torch.equal(new, old)):
				break
			if float(stats[0].item()) <= tol:
				break
		# E-step on the final centres: the labels sklearn returns (after a strict stop the centres did not move, so
		# these ARE the labels of the last iteration) and their inertia (_kmeans_single_lloyd, :740-758)
		final = lab[0]
		e._call("cs_lloyd_step_rows64", self.rows.data_ptr(), n, c[cur].data_ptr(), K, final.data_ptr(), inert.data_ptr())
		return final, float(inert.item())

	def _relocate_host(self, labels, c_old, sums, counts):
		"""_relocate_empty_clusters_dense (sklearn/cluster/_k_means_common.pyx:167-211) from the device labels."""
		torch = _torch()
		X = self.rows.cpu().numpy()
		lab = labels.cpu().numpy()
		C = c_old.cpu().numpy()
		s, w = sums.cpu().numpy().copy(), counts.cpu().numpy().copy()
		empty = np.nonzero(w == 0)[0]
		dist = ((X - C[lab]) ** 2).sum(axis=1)
		# sklearn returns when np.max(distances) == 0.  With fewer DISTINCT rows than clusters every point sits on its
		# centre up to the rounding of sum / count (1e-32 here), and whether scikit-learn relocates then depends on
		# its summation-order noise; distances at that level count as zero here (documented, DESIGN.md deviation 9)
		if dist.max() <= 1e-24 * max(1.0, float((X * X).sum(axis=1).max())):
			return
		order = np.lexsort((np.arange(len(dist)), -dist))[:len(empty)]
		for new_id, far in zip(empty, order):
			old_id = lab[far]
			s[old_id] -= X[far]
			s[new_id] = X[far]
			w[new_id] = 1.0
			w[old_id] -= 1.0
		sums.copy_(torch.from_numpy(s))
		counts.copy_(torch.from_numpy(w))

	def fit_predict(self, K: int, n_init: int = 10, max_iter: int = 300, seed: int = 42) -> np.ndarray:
		torch = _torch()
		# _tolerance: mean of the per-feature variances x 1e-4 (sklearn/cluster/_kmeans.py:285-293)
		tol = float(torch.var(self.rows, dim=0, unbiased=False).mean().item()) * 1e-4
		_, inits = self.eng.kmeanspp_seeds(None, None, K, n_init, seed, rows=self.rows)
		best = None
		for init in inits:
			lab, inertia = self._run(init, max_iter, tol)
			if best is None or (inertia < best[1] and not self._same_clustering(lab, best[0], K)):
				best = (lab.clone(), inertia)
		return best[0].cpu().numpy().astype(np.int64)

	@staticmethod
	def _same_clustering(l1, l2, K: int) -> bool:
		"""_is_same_clustering (sklearn/cluster/_k_means_common.pyx:314-328) on the device labels."""
		torch = _torch()
		pair = torch.unique(l1.to(torch.int64) * K + l2.to(torch.int64))
		return bool(torch.unique(pair // K).numel() == pair.numel())


class SampleKMeans:
	"""KMeans(n_clusters=K, random_state=42, n_init=10, max_iter=max_iter).fit(X) for a SAMPLE of n <= 5120 fp64 rows —
	the palette fit of simplify_colors_perceptual_fast (reference color_simplify.py:669-675) — on the device:
	the rows are mean-centred as KMeans.fit does (sklearn/cluster/_kmeans.py:1487-1494), the n_init k-means++
	seedings run in lockstep (Engine.kmeanspp_seeds on rows), then ONE launch runs the n_init complete Lloyd loops
	side by side, one CTA per initialisation with the rows in shared memory (cs_kmeans_fit_rows64_small).  The host
	keeps the run of least inertia exactly as KMeans.fit does (:1536-1541, _is_same_clustering) and adds the mean
	back.  Distances are the direct fp64 formula (scikit-learn: |x|^2 - 2 x.c + |c|^2 through GEMM), sums run in a
	fixed order: centres agree with scikit-learn's to rounding (1e-12 relative; scikit-learn's own depend on its
	thread count at that level), labels except on ties at rounding level."""

	MAX_ROWS = 5120

	def __init__(self, eng: Engine):
		self.eng = eng
		self.cluster_centers_ = None
		self.labels_ = None
		self.inertia_ = None
		self.n_iter_ = None

	def fit(self, X: np.ndarray, K: int, n_init: int = 10, max_iter: int = 300, seed: int = 42, tol: float = 1e-4) -> "SampleKMeans":
		torch = _torch()
		e = self.eng
		X = np.ascontiguousarray(X, dtype=np.float64)
		n = int(X.shape[0])
		if not 1 <= n <= self.MAX_ROWS or X.shape[1] != 3:
			raise ValueError(f"SampleKMeans takes between 1 and {self.MAX_ROWS} rows of 3 features")
		tol_abs = float(np.mean(np.var(X, axis=0)) * tol)  # _tolerance (:285-293), on the rows as given
		mean = X.mean(axis=0)
		Xc = X - mean
		d_rows = torch.from_numpy(Xc).to(e.dev)
		_, inits = e.kmeanspp_seeds(None, None, K, n_init, seed, rows=d_rows)
		d_init = torch.from_numpy(np.ascontiguousarray(np.stack(inits), dtype=np.float64)).to(e.dev)
		d_cent = e.empty((n_init, K, 3), torch.float64)
		d_lab = e.empty((n_init, n), torch.uint8)
		d_stats = e.empty((n_init, 4), torch.float64)
		e._call("cs_kmeans_fit_rows64_small", d_rows.data_ptr(), n, d_init.data_ptr(), n_init, K, int(max_iter), tol_abs,
		        d_cent.data_ptr(), d_lab.data_ptr(), d_stats.data_ptr())
		stats = d_stats.cpu().numpy()
		labels = d_lab.cpu().numpy()
		best = None
		for i in range(n_init):
			inertia = float(stats[i, 0])
			if best is None or (inertia < float(stats[best, 0]) and not _same_clustering_np(labels[i], labels[best], K)):
				best = i
		self.cluster_centers_ = d_cent[best].cpu().numpy() + mean
		self.labels_ = labels[best].astype(np.int32)
		self.inertia_ = float(stats[best, 0])
		self.n_iter_ = int(stats[best, 1])
		return self


def _same_clustering_np(l1: np.ndarray, l2: np.ndarray, K: int) -> bool:
	"""_is_same_clustering (sklearn/cluster/_k_means_common.pyx:314-328), as written there: a one-way mapping test."""
	pair = np.unique(l1.astype(np.int64) * K + l2.astype(np.int64))
	return bool(np.unique(pair // K).size == pair.size)  # every id of l1 maps to ONE id of l2


_engines: dict[int, Engine] = {}


def get_engine(device: int | None = None) -> Engine:
	torch = _torch()
	if not torch.cuda.is_available():
		raise _ffi.ColorSimplifyError(
			"no CUDA device: image_segmenter_b200 runs on B200 (sm_100a) only and has no CPU fallback")
	if device is None:
		device = torch.cuda.current_device()
	e = _engines.get(device)
	if e is None:
		e = _engines[device] = Engine(device)
	return e
