// ccl.cu — connected components of equal-colour opaque pixels in ONE pass set, sm_100a.
//
// Replaces the per-colour loop of analyze_regions (app/processing/region_cleanup.py:9-130): for every
// unique colour the reference builds a binary mask and calls cv.connectedComponentsWithStats
// (O(K * N)); the immediate consumer of the colour-simplified image.  Here a
// single label-equivalence (union-find) labelling keyed on "same RGB, alpha > 0, 4- or 8-connected"
// labels all colours at once; areas, bounding boxes and the quantity that fixes OpenCV's label order
// come from one more pass, and per-colour label images are produced on demand.
//
//   ccl_init / ccl_merge / ccl_compress   label[i] = smallest linear index of i's component (-1 = transparent)
//   ccl_stats                             per component: area, bbox, order key
//   ccl_extract                           the `labels` / `color_mask` arrays the reference returns per colour
//
// OpenCV numbers the components of a mask in the order its scan first meets them: raster order of the
// first pixel for 4-connectivity (SAUF), raster order of the first 2x2 BLOCK for 8-connectivity
// (block-based decision trees) — verified against cv2 4.13 in tests/test_oracle_regions.py.  The order
// key computed here is exactly that (min over the component's pixels of the block / pixel index).
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool same_region(uint32_t a, uint32_t b) {
	// both opaque and equal RGB (alpha values may differ: the reference compares rgb and alpha > 0 only)
	return (a >> 24) && (b >> 24) && ((a ^ b) & 0x00FFFFFFu) == 0u;
}

__device__ __forceinline__ int find_root(const int *L, int i) {
	const volatile int *V = L;  // other threads re-parent nodes concurrently: always re-read
	int p = V[i];
	while (p != i) { i = p; p = V[i]; }
	return i;
}

// link the larger root under the smaller one (so a component's root ends up being its smallest index)
__device__ __forceinline__ void unite(int *L, int a, int b) {
	while (true) {
		a = find_root(L, a);
		b = find_root(L, b);
		if (a == b) return;
		if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b
		const int old = atomicMin(&L[a], b);
		if (old == a) return;
		a = old;  // somebody re-parented a meanwhile: retry from there
	}
}

__global__ void __launch_bounds__(kThreads) ccl_init_kernel(const uint32_t *__restrict__ px, long long n, int *L) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
		L[i] = (px[i] >> 24) ? (int)i : -1;
}

__global__ void __launch_bounds__(kThreads) ccl_merge_kernel(const uint32_t *__restrict__ px, int W, int H, int conn8,
                                                             int *L) {
	const long long n = (long long)W * H;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t c = px[i];
		if (!(c >> 24)) continue;
		const int x = (int)(i % W), y = (int)(i / W);
		// backward neighbours only: W, and N (+ NW, NE for 8-connectivity)
		if (x > 0 && same_region(c, px[i - 1])) unite(L, (int)i, (int)(i - 1));
		if (y > 0) {
			if (same_region(c, px[i - W])) unite(L, (int)i, (int)(i - W));
			if (conn8) {
				if (x > 0 && same_region(c, px[i - W - 1])) unite(L, (int)i, (int)(i - W - 1));
				if (x + 1 < W && same_region(c, px[i - W + 1])) unite(L, (int)i, (int)(i - W + 1));
			}
		}
	}
}

__global__ void __launch_bounds__(kThreads) ccl_compress_kernel(long long n, int *L) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
		if (L[i] >= 0) L[i] = find_root(L, (int)i);
}

// roots (L[i] == i) in raster order: count per tile, scan (select.cu's scan kernel is reused through the
// host wrapper), scatter; rank[i] = component id at root pixels
constexpr int kTile = 4096;
__global__ void __launch_bounds__(kThreads) ccl_root_count_kernel(const int *__restrict__ L, long long n,
                                                                  unsigned long long *tile_counts) {
	const long long base = (long long)blockIdx.x * kTile;
	uint32_t c = 0;
	for (int j = threadIdx.x; j < kTile; j += kThreads)
		if (base + j < n) c += L[base + j] == (int)(base + j);
	__shared__ uint32_t s[kThreads / 32];
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < kThreads / 32; ++w) t += s[w];
		tile_counts[blockIdx.x] = t;
	}
}
__global__ void __launch_bounds__(1024) ccl_scan_kernel(unsigned long long *v, long long m, unsigned long long *total) {
	__shared__ unsigned long long s[1024];
	__shared__ unsigned long long carry;
	if (threadIdx.x == 0) carry = 0ull;
	__syncthreads();
	for (long long t0 = 0; t0 < m; t0 += 1024) {
		const long long i = t0 + threadIdx.x;
		const unsigned long long x = i < m ? v[i] : 0ull;
		s[threadIdx.x] = x;
		__syncthreads();
		for (int o = 1; o < 1024; o <<= 1) {
			const unsigned long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0ull;
			__syncthreads();
			s[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < m) v[i] = carry + s[threadIdx.x] - x;
		__syncthreads();
		if (threadIdx.x == 1023) carry += s[1023];
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(kThreads) ccl_root_scatter_kernel(const int *__restrict__ L, long long n,
                                                                    const unsigned long long *__restrict__ tile_off,
                                                                    int *__restrict__ rank, int *__restrict__ roots,
                                                                    long long capacity) {
	const long long base = (long long)blockIdx.x * kTile;
	__shared__ uint32_t warp_cnt[kThreads / 32];
	__shared__ unsigned long long running;
	if (threadIdx.x == 0) running = tile_off[blockIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int j0 = 0; j0 < kTile; j0 += kThreads) {
		const long long i = base + j0 + threadIdx.x;
		const bool is_root = i < n && L[i] == (int)i;
		const uint32_t m = __ballot_sync(0xffffffffu, is_root);
		if (lane == 0) warp_cnt[w] = __popc(m);
		__syncthreads();
		unsigned long long off = running;
		for (int q = 0; q < w; ++q) off += warp_cnt[q];
		if (i < n) {
			int r = -1;
			if (is_root) {
				const long long pos = (long long)off + __popc(m & ((1u << lane) - 1u));
				r = (int)pos;
				if (pos < capacity) roots[pos] = (int)i;
			}
			rank[i] = r;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (int q = 0; q < kThreads / 32; ++q) t += warp_cnt[q];
			running += t;
		}
		__syncthreads();
	}
}

// per component: area, bbox {minx, miny, maxx, maxy}, order key.  One thread walks a horizontal run of
// kRun pixels and flushes its local aggregate whenever the component changes, so a large region costs a
// handful of atomics per run instead of one per pixel.
constexpr int kRun = 32;
__global__ void __launch_bounds__(kThreads) ccl_stats_kernel(const int *__restrict__ L, const int *__restrict__ rank, int W,
                                                             int H, int conn8, uint32_t *area, int *bbox,
                                                             unsigned long long *order_key) {
	const int runs_per_row = (W + kRun - 1) / kRun;
	const long long total = (long long)runs_per_row * H;
	const long long stride = (long long)gridDim.x * kThreads;
	const unsigned long long W2 = (unsigned long long)((W + 1) / 2);
	for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += stride) {
		const int y = (int)(t / runs_per_row), x0 = (int)(t % runs_per_row) * kRun;
		const int x1 = x0 + kRun < W ? x0 + kRun : W;
		int cur = -1, cnt = 0, xa = 0, xb = 0;
		auto flush = [&]() {
			if (cur < 0) return;
			atomicAdd(area + cur, (uint32_t)cnt);
			atomicMin(bbox + 4 * cur + 0, xa); atomicMin(bbox + 4 * cur + 1, y);
			atomicMax(bbox + 4 * cur + 2, xb); atomicMax(bbox + 4 * cur + 3, y);
			// first 2x2 block (8-connectivity) or first pixel (4-connectivity) of the run, in raster order
			const unsigned long long key = conn8 ? (unsigned long long)(y >> 1) * W2 + (unsigned long long)(xa >> 1)
			                                     : (unsigned long long)y * (unsigned long long)W + (unsigned long long)xa;
			atomicMin(order_key + cur, key);
		};
		for (int x = x0; x < x1; ++x) {
			const int root = L[(long long)y * W + x];
			const int r = root >= 0 ? rank[root] : -1;
			if (r != cur) {
				flush();
				cur = r; cnt = 0; xa = x;
			}
			if (r >= 0) { ++cnt; xb = x; }
		}
		flush();
	}
}

// the per-colour arrays of the reference: labels (int32: component number within the colour, 0 elsewhere)
// and color_mask (u8: 255 where the pixel has the colour and is opaque)
__global__ void __launch_bounds__(kThreads) ccl_extract_kernel(const int *__restrict__ L, const int *__restrict__ rank,
                                                               long long n, const int *__restrict__ comp_color,
                                                               const int *__restrict__ comp_local, int color,
                                                               int *__restrict__ out_labels, uint8_t *__restrict__ out_mask) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const int root = L[i];
		int lab = 0;
		uint8_t m = 0;
		if (root >= 0) {
			const int r = rank[root];
			if (comp_color[r] == color) { lab = comp_local[r]; m = 255; }
		}
		if (out_labels) out_labels[i] = lab;
		if (out_mask) out_mask[i] = m;
	}
}

__global__ void __launch_bounds__(kThreads) ccl_fill_kernel(int n_comp, int *bbox, unsigned long long *order_key,
                                                            uint32_t *area) {
	const int i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= n_comp) return;
	area[i] = 0u;
	bbox[4 * i] = 0x7fffffff; bbox[4 * i + 1] = 0x7fffffff; bbox[4 * i + 2] = -1; bbox[4 * i + 3] = -1;
	order_key[i] = ~0ull;
}

} // namespace
} // namespace cs

using namespace cs;

#define CS_STREAM ((cudaStream_t)stream)

extern "C" int cs_ccl_label(cs_ctx *ctx, const uint8_t *d_rgba, int width, int height, int connectivity,
                            int32_t *d_labels, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_labels, "null pointer");
	CS_REQUIRE(width > 0 && height > 0 && (long long)width * height < (1LL << 31), "image must have 1 .. 2^31-1 pixels");
	CS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
	const long long n = (long long)width * height;
	const uint32_t *px = reinterpret_cast<const uint32_t *>(d_rgba);
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	ccl_init_kernel<<<grid, kThreads, 0, CS_STREAM>>>(px, n, d_labels);
	ccl_merge_kernel<<<grid, kThreads, 0, CS_STREAM>>>(px, width, height, connectivity == 8, d_labels);
	ccl_compress_kernel<<<grid, kThreads, 0, CS_STREAM>>>(n, d_labels);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_ccl_roots(cs_ctx *ctx, const int32_t *d_labels, int64_t n, int32_t *d_rank, int32_t *d_roots,
                            int64_t capacity, unsigned long long *d_count, void *stream) {
	CS_REQUIRE(ctx && d_labels && d_count, "null pointer");
	CS_REQUIRE(n > 0 && n < (1LL << 31) && capacity >= 0, "bad n or capacity");
	const long long ntiles = (n + kTile - 1) / kTile;
	CS_REQUIRE(ntiles <= (long long)kMaxPartialBlocks * kMaxPartialVals, "n too large for the scan scratch");
	unsigned long long *tiles = reinterpret_cast<unsigned long long *>(ctx->d_partials);
	ccl_root_count_kernel<<<(unsigned)ntiles, kThreads, 0, CS_STREAM>>>(d_labels, n, tiles);
	ccl_scan_kernel<<<1, 1024, 0, CS_STREAM>>>(tiles, ntiles, d_count);
	if (d_rank)
		ccl_root_scatter_kernel<<<(unsigned)ntiles, kThreads, 0, CS_STREAM>>>(d_labels, n, tiles, d_rank, d_roots,
		                                                                      d_roots ? capacity : 0);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_ccl_stats(cs_ctx *ctx, const int32_t *d_labels, const int32_t *d_rank, int width, int height,
                            int connectivity, int n_comp, uint32_t *d_area, int32_t *d_bbox,
                            unsigned long long *d_order_key, void *stream) {
	CS_REQUIRE(ctx && d_labels && d_rank && d_area && d_bbox && d_order_key, "null pointer");
	CS_REQUIRE(width > 0 && height > 0 && n_comp >= 0, "bad size");
	CS_REQUIRE(connectivity == 4 || connectivity == 8, "connectivity must be 4 or 8");
	if (n_comp == 0) return 0;
	ccl_fill_kernel<<<(n_comp + kThreads - 1) / kThreads, kThreads, 0, CS_STREAM>>>(n_comp, d_bbox, d_order_key, d_area);
	const long long runs = (long long)((width + kRun - 1) / kRun) * height;
	ccl_stats_kernel<<<grid_for(ctx, (runs + kThreads - 1) / kThreads, 8), kThreads, 0, CS_STREAM>>>(
	    d_labels, d_rank, width, height, connectivity == 8, d_area, d_bbox, d_order_key);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_ccl_extract(cs_ctx *ctx, const int32_t *d_labels, const int32_t *d_rank, int64_t n,
                              const int32_t *d_comp_color, const int32_t *d_comp_local, int color,
                              int32_t *d_out_labels, uint8_t *d_out_mask, void *stream) {
	CS_REQUIRE(ctx && d_labels && d_rank && d_comp_color && d_comp_local, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	if (n == 0) return 0;
	ccl_extract_kernel<<<grid_for(ctx, (n + kThreads - 1) / kThreads, 8), kThreads, 0, CS_STREAM>>>(
	    d_labels, d_rank, n, d_comp_color, d_comp_local, color, d_out_labels, d_out_mask);
	CS_CUDA(cudaGetLastError());
	return 0;
}
