"""Pillow's MEDIANCUT quantiser restated in NumPy / pure Python (integer, bit-exact).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference calls
`Image.quantize(colors=K, method=Image.Quantize.MEDIANCUT)` at
app/processing/color_simplify.py:145 (median_cut) and :201 ("octree" — which also asks for
MEDIANCUT).  Pillow (requirements.txt: `Pillow>=10.0`, unpinned; image has 12.2.0) implements
it in src/libImaging/Quant.c + QuantHeap.c, compiled into PIL/_imaging*.so — the C sources are
not on disk.  This file restates the published algorithm (function names of Quant.c in the
docstrings) and is pinned by tests/test_oracle_mediancut.py against Pillow itself: palettes and
index maps bit-identical on generic, equal-population and tie-heavy images.
"""
from __future__ import annotations

import numpy as np

MAX_HASH_ENTRIES = 65536


def histogram_cells(rgb: np.ndarray):
	"""create_pixel_hash: count pixels per (r>>s, g>>s, b>>s) cell, s = the smallest shift with
	<= 65536 distinct cells (Pillow bumps its scale whenever the hash table grows past 65536
	entries while scanning, which ends at the same s).  Returns (shift, cells (n,3) uint8, counts)."""
	rgb = np.ascontiguousarray(rgb.reshape(-1, 3))
	key = (rgb[:, 0].astype(np.uint32) << 16) | (rgb[:, 1].astype(np.uint32) << 8) | rgb[:, 2].astype(np.uint32)
	full_keys, full_counts = np.unique(key, return_counts=True)
	for shift in range(8):
		bits = 8 - shift
		r = (full_keys >> 16) >> shift
		g = ((full_keys >> 8) & 0xFF) >> shift
		b = (full_keys & 0xFF) >> shift
		ck = (r << (2 * bits)) | (g << bits) | b
		cells, inv = np.unique(ck, return_inverse=True)
		if len(cells) <= MAX_HASH_ENTRIES:
			counts = np.bincount(inv, weights=full_counts).astype(np.int64)
			mask = (1 << bits) - 1
			cell_rgb = np.stack([(cells >> (2 * bits)) & mask, (cells >> bits) & mask, cells & mask], axis=1)
			return shift, cell_rgb.astype(np.uint8), counts, cells
	raise AssertionError("unreachable: shift 7 has at most 8 cells")


class _Heap:
	"""ImagingQuantHeapAdd / ImagingQuantHeapRemove (QuantHeap.c): 1-indexed array max-heap."""

	def __init__(self, key):
		self.h = [None]
		self.key = key

	def add(self, v):
		self.h.append(v)
		k = len(self.h) - 1
		while k != 1:
			if self.key(v) - self.key(self.h[k // 2]) <= 0:
				break
			self.h[k] = self.h[k // 2]
			k //= 2
		self.h[k] = v

	def remove(self):
		n = len(self.h) - 1
		if n == 0:
			return None
		r = self.h[1]
		v = self.h.pop()
		n -= 1
		k = 1
		while k * 2 <= n:
			l = k * 2
			if l < n and self.key(self.h[l]) - self.key(self.h[l + 1]) < 0:
				l += 1
			if self.key(v) - self.key(self.h[l]) > 0:
				break
			self.h[k] = self.h[l]
			k = l
		if n >= 1:
			self.h[k] = v
		return r


class _Box:
	__slots__ = ("idx", "count", "l", "r")

	def __init__(self, idx, count):
		self.idx, self.count, self.l, self.r = idx, int(count), None, None


def _split(box: _Box, cells: np.ndarray, counts: np.ndarray):
	"""split() + splitlists(): axis = first max of (dr*77, dg*150, db*29); walk the cells in
	descending axis value until 2*acc > count, extend over equal values -> left/high child; an
	empty right child receives the cells holding the minimum axis value."""
	c = cells[box.idx].astype(np.int64)
	ext = c.max(axis=0) - c.min(axis=0)
	f = ext * np.array([77, 150, 29])
	axis = 0
	best = f[0]
	for a in (1, 2):
		if best < f[a]:
			best, axis = f[a], a
	order = np.argsort(-c[:, axis], kind="stable")
	vals = c[order, axis]
	cnts = counts[box.idx][order]
	acc = np.cumsum(cnts)
	pos = int(np.argmax(acc * 2 > box.count)) + 1  # first position where the walk breaks
	if pos < len(order):
		split_val = vals[pos - 1]
		while pos < len(order) and vals[pos] == split_val:
			pos += 1
	if pos == len(order):
		tail_val = vals[-1]
		while pos > 0 and vals[pos - 1] == tail_val:
			pos -= 1
	left = _Box(box.idx[order[:pos]], cnts[:pos].sum())
	right = _Box(box.idx[order[pos:]], cnts[pos:].sum())
	box.l, box.r = left, right
	box.idx = None


def median_cut_boxes(cells: np.ndarray, counts: np.ndarray, n_colors: int) -> np.ndarray:
	"""median_cut() + annotate_hash_table(): returns the palette index of every cell
	(depth-first leaf order, left child first).  A popped box of volume 1 is dropped."""
	root = _Box(np.arange(len(cells)), counts.sum())
	heap = _Heap(lambda b: b.count)
	heap.add(root)
	for _ in range(n_colors - 1):
		while True:
			node = heap.remove()
			if node is None:
				break
			c = cells[node.idx].astype(np.int64)
			vol = int(np.prod(c.max(axis=0) - c.min(axis=0) + 1))
			if vol != 1:
				break
		if node is None:
			break
		_split(node, cells, counts)
		heap.add(node.l)
		heap.add(node.r)
	cell_box = np.zeros(len(cells), dtype=np.int64)
	nxt = 0
	stack = [root]
	while stack:
		b = stack.pop()
		if b.l is not None:
			stack.append(b.r)
			stack.append(b.l)
			continue
		cell_box[b.idx] = nxt
		nxt += 1
	return cell_box


def box_palette(rgb_keys: np.ndarray, key_counts: np.ndarray, key_box: np.ndarray, n_boxes: int) -> np.ndarray:
	"""compute_palette_from_median_cut: per box, channel = (int)(0.5 + sum / count) over the
	UNSCALED pixels, with the sums and counts held in uint32 as in the C code."""
	pal = np.zeros((n_boxes, 3), dtype=np.uint8)
	ch = [(rgb_keys >> 16) & 0xFF, (rgb_keys >> 8) & 0xFF, rgb_keys & 0xFF]
	cnt = np.bincount(key_box, weights=key_counts, minlength=n_boxes).astype(np.uint64) & 0xFFFFFFFF
	for j in range(3):
		s = np.zeros(n_boxes, dtype=np.uint64)
		np.add.at(s, key_box, (ch[j].astype(np.uint64) * key_counts.astype(np.uint64)))
		s &= 0xFFFFFFFF
		pal[:, j] = (0.5 + s.astype(np.float64) / cnt.astype(np.float64)).astype(np.int64).astype(np.uint8)
	return pal


def map_to_palette(rgb_unique: np.ndarray, own: np.ndarray, palette: np.ndarray) -> np.ndarray:
	"""map_image_pixels_from_median_box: start at the own box's entry, visit entries in ascending
	(squared palette distance from the own entry, index), replace only on a strictly smaller
	squared distance.  Equivalent closed form used here: if the own entry attains the minimum it
	wins, otherwise the minimiser with the smallest (palette distance from own, index)."""
	P = palette.astype(np.int64)
	n_pal = len(P)
	pd = ((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)  # avgDist
	rank = np.empty((n_pal, n_pal), dtype=np.int64)  # position of j in own's sorted scan order
	for i in range(n_pal):
		order = np.lexsort((np.arange(n_pal), pd[i]))
		rank[i, order] = np.arange(n_pal)
	out = np.empty(len(rgb_unique), dtype=np.int64)
	step = 1 << 14
	X = rgb_unique.astype(np.int64)
	for s in range(0, len(X), step):
		x = X[s:s + step]
		o = own[s:s + step]
		d = ((x[:, None, :] - P[None, :, :]) ** 2).sum(-1)
		tie_key = rank[o] + 1
		tie_key[np.arange(len(o)), o] = 0  # own first
		out[s:s + step] = np.argmin(d * (n_pal + 1) + tie_key, axis=1)
	return out


def quantize(rgb: np.ndarray, n_colors: int):
	"""Image.quantize(colors=n_colors, method=MEDIANCUT, kmeans=0) on an (H,W,3) or (N,3) uint8
	array.  Returns (palette (P,3) uint8 with P <= n_colors, indices, shape like rgb[...,0])."""
	shape = rgb.shape[:-1]
	flat = np.ascontiguousarray(rgb.reshape(-1, 3))
	shift, cells, counts, cell_keys = histogram_cells(flat)
	cell_box = median_cut_boxes(cells, counts, n_colors)
	n_boxes = int(cell_box.max()) + 1
	key = (flat[:, 0].astype(np.uint32) << 16) | (flat[:, 1].astype(np.uint32) << 8) | flat[:, 2].astype(np.uint32)
	ukeys, inv, ucounts = np.unique(key, return_inverse=True, return_counts=True)
	bits = 8 - shift
	uck = ((((ukeys >> 16) & 0xFF) >> shift) << (2 * bits)) | ((((ukeys >> 8) & 0xFF) >> shift) << bits) | ((ukeys & 0xFF) >> shift)
	ubox = cell_box[np.searchsorted(cell_keys, uck)]
	pal = box_palette(ukeys, ucounts, ubox, n_boxes)
	urgb = np.stack([(ukeys >> 16) & 0xFF, (ukeys >> 8) & 0xFF, ukeys & 0xFF], axis=1)
	uidx = map_to_palette(urgb, ubox, pal)
	return pal, uidx[inv].reshape(shape).astype(np.uint8)
