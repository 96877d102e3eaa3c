// mediancut.cpp — host side of the median-cut quantiser: the box tree over the (<= 65536-cell)
// colour histogram.  Replaces Pillow's median_cut()/split()/splitlists() and the array heap of
// QuantHeap.c (compiled into PIL/_imaging; reached from Image.quantize(method=MEDIANCUT) at
// app/processing/color_simplify.py:145 and :201).  The per-pixel passes around it (histogram,
// box means, nearest-palette map) are CUDA kernels in hist.cu; this step is sequential by
// nature (K-1 dependent splits) and touches at most 65536 cells, so it runs on the host.
//
// Semantics reproduced (verified bit-for-bit against Pillow 12.2.0 in tests/test_oracle_mediancut.py
// through the oracle, and against this file in tests/test_abi.py):
//   * boxes are popped from a 1-indexed array max-heap keyed on pixel count, Pillow's exact
//     sift rules (ties depend on insertion history);
//   * a popped box whose cell-extent volume is 1 is dropped (stays a leaf); the loop runs
//     n_colors-1 times or until the heap is empty;
//   * split axis = first maximum of (dr*77, dg*150, db*29) on the scaled cell values;
//   * cells are walked in DESCENDING axis value accumulating counts until 2*acc > count, the
//     walk is extended over cells sharing the last value; that is the left/high child.  If the
//     right child would be empty the cells holding the minimum axis value move to it;
//   * palette order = depth-first leaves, left (high) child first.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/colorsimplify.h"

namespace cs {
void set_error(const char *fmt, ...);
}

namespace {

struct Box {
	std::vector<uint32_t> cells;  // indices into the cell arrays
	uint32_t pixel_count = 0;
	int left = -1, right = -1;
	int volume = -1;
};

struct Cells {
	const uint8_t *v[3];  // scaled r, g, b per cell
	const uint32_t *count;
};

int box_volume(Box &b, const Cells &c) {
	if (b.volume >= 0) return b.volume;
	if (b.cells.empty()) return b.volume = 0;
	int lo[3] = {255, 255, 255}, hi[3] = {0, 0, 0};
	for (uint32_t i : b.cells)
		for (int a = 0; a < 3; ++a) {
			lo[a] = std::min<int>(lo[a], c.v[a][i]);
			hi[a] = std::max<int>(hi[a], c.v[a][i]);
		}
	return b.volume = (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
}

// Pillow's heap comparison: (int)A->pixelCount - (int)B->pixelCount
inline int cmp_boxes(const std::vector<Box> &boxes, int a, int b) {
	return (int)boxes[a].pixel_count - (int)boxes[b].pixel_count;
}

struct QuantHeap {  // ImagingQuantHeapAdd / ImagingQuantHeapRemove
	std::vector<int> h{0};  // slot 0 unused
	void add(const std::vector<Box> &boxes, int val) {
		h.push_back(val);
		size_t k = h.size() - 1;
		while (k != 1) {
			if (cmp_boxes(boxes, val, h[k / 2]) <= 0) break;
			h[k] = h[k / 2];
			k >>= 1;
		}
		h[k] = val;
	}
	bool remove(const std::vector<Box> &boxes, int &out) {
		size_t n = h.size() - 1;
		if (n == 0) return false;
		out = h[1];
		const int v = h[n];
		h.pop_back();
		--n;
		size_t k = 1, l;
		for (; k * 2 <= n; k = l) {
			l = k * 2;
			if (l < n && cmp_boxes(boxes, h[l], h[l + 1]) < 0) ++l;
			if (cmp_boxes(boxes, v, h[l]) > 0) break;
			h[k] = h[l];
		}
		if (n >= 1) h[k] = v;
		return true;
	}
};

void split_box(std::vector<Box> &boxes, int idx, const Cells &c) {
	int lo[3] = {255, 255, 255}, hi[3] = {0, 0, 0};
	for (uint32_t i : boxes[idx].cells)
		for (int a = 0; a < 3; ++a) {
			lo[a] = std::min<int>(lo[a], c.v[a][i]);
			hi[a] = std::max<int>(hi[a], c.v[a][i]);
		}
	const int f[3] = {(hi[0] - lo[0]) * 77, (hi[1] - lo[1]) * 150, (hi[2] - lo[2]) * 29};
	int axis = 0, best = f[0];
	for (int a = 1; a < 3; ++a)
		if (best < f[a]) { best = f[a]; axis = a; }

	// descending counting sort of the box's cells by the axis value
	const std::vector<uint32_t> &src = boxes[idx].cells;
	uint32_t start[257] = {0};
	for (uint32_t i : src) ++start[255 - c.v[axis][i] + 1];
	for (int v = 0; v < 256; ++v) start[v + 1] += start[v];
	std::vector<uint32_t> order(src.size());
	for (uint32_t i : src) order[start[255 - c.v[axis][i]]++] = i;

	const uint32_t pixel_count = boxes[idx].pixel_count;
	size_t pos = 0;
	uint32_t acc = 0, n_left = 0;
	while (pos < order.size()) {
		acc += c.count[order[pos]];
		n_left += c.count[order[pos]];
		++pos;
		if (acc * 2u > pixel_count) break;  // uint32 arithmetic, as in splitlists()
	}
	if (pos < order.size()) {
		const int split_val = c.v[axis][order[pos - 1]];
		while (pos < order.size() && c.v[axis][order[pos]] == split_val) {
			n_left += c.count[order[pos]];
			++pos;
		}
	}
	uint32_t n_right = 0;
	for (size_t j = pos; j < order.size(); ++j) n_right += c.count[order[j]];
	if (n_right == 0) {
		const int tail_val = c.v[axis][order.back()];
		while (pos > 0 && c.v[axis][order[pos - 1]] == tail_val) {
			--pos;
			n_left -= c.count[order[pos]];
			n_right += c.count[order[pos]];
		}
	}
	Box l, r;
	l.cells.assign(order.begin(), order.begin() + pos);
	r.cells.assign(order.begin() + pos, order.end());
	l.pixel_count = n_left;
	r.pixel_count = n_right;
	const int li = (int)boxes.size();
	boxes.push_back(std::move(l));
	boxes.push_back(std::move(r));
	boxes[idx].left = li;
	boxes[idx].right = li + 1;
	boxes[idx].cells.clear();
	boxes[idx].cells.shrink_to_fit();
}

} // namespace

extern "C" int cs_median_cut_boxes(const uint32_t *h_keys, const uint32_t *h_counts, uint32_t n, int shift,
                                   int n_colors, uint16_t *h_cell_box, int *n_boxes) {
	if (!h_keys || !h_counts || !h_cell_box || !n_boxes || shift < 0 || shift > 7 || n_colors < 1 || n == 0) {
		cs::set_error("cs_median_cut_boxes: bad argument");
		return CS_ERR_ARG;
	}
	const int bits = 8 - shift;
	const uint32_t mask = (1u << bits) - 1u;
	std::vector<uint8_t> r(n), g(n), b(n);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		r[i] = (uint8_t)((h_keys[i] >> (2 * bits)) & mask);
		g[i] = (uint8_t)((h_keys[i] >> bits) & mask);
		b[i] = (uint8_t)(h_keys[i] & mask);
		total += h_counts[i];
	}
	Cells c{{r.data(), g.data(), b.data()}, h_counts};
	std::vector<Box> boxes;
	boxes.reserve(2 * (size_t)n_colors + 2);
	Box root;
	root.cells.resize(n);
	for (uint32_t i = 0; i < n; ++i) root.cells[i] = i;
	root.pixel_count = (uint32_t)total;
	boxes.push_back(std::move(root));
	QuantHeap heap;
	heap.add(boxes, 0);
	int remaining = n_colors;
	bool done = false;
	while (--remaining > 0 && !done) {
		int cur;
		for (;;) {
			if (!heap.remove(boxes, cur)) { done = true; break; }
			if (box_volume(boxes[cur], c) != 1) break;
		}
		if (done) break;
		split_box(boxes, cur, c);
		heap.add(boxes, boxes[cur].left);
		heap.add(boxes, boxes[cur].right);
	}
	// palette index = depth-first leaf order, left child first
	int next = 0;
	std::vector<int> stack{0};
	while (!stack.empty()) {
		const int i = stack.back();
		stack.pop_back();
		if (boxes[i].left >= 0) {
			stack.push_back(boxes[i].right);
			stack.push_back(boxes[i].left);
			continue;
		}
		if (boxes[i].cells.empty()) continue;
		if (next >= 65535) {
			cs::set_error("cs_median_cut_boxes: too many boxes");
			return CS_ERR_ARG;
		}
		for (uint32_t cell : boxes[i].cells) h_cell_box[cell] = (uint16_t)next;
		++next;
	}
	*n_boxes = next;
	return 0;
}
