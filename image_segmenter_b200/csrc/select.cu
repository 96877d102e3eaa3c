// select.cu — the masking / sampling plumbing around the clustering kernels, sm_100a.
//
// The reference builds its working sets with NumPy boolean indexing on the host:
//   rgb[non_transparent], rgb_flat[non_black_mask]          (color_simplify.py:49-66, 439-466,
//                                                            599-655, 946-967)
//   rgb_flat[np.random.choice(len(rgb_flat), m, False)]     (:443-445, :633-635)
//   np.var / np.mean of the filtered feature rows           (sklearn/cluster/_kmeans.py:285-293,
//                                                            1487-1490 via KMeans.fit)
//   _is_same_clustering                                     (sklearn/cluster/_k_means_common.pyx:314-328)
// These kernels do the same selections on the device, order-preserving, so that the i-th
// selected pixel here is the i-th row of the reference's compacted array.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;
constexpr int kTile = 4096;  // pixels per block in the compaction passes

__device__ __forceinline__ bool selected(uint32_t w, int mask_mode, int min_bright) {
	return px_selected(w, mask_mode, min_bright);
}

__global__ void __launch_bounds__(kThreads) select_count_kernel(const uint32_t *__restrict__ px, long long n,
                                                                int mask_mode, int min_bright,
                                                                unsigned long long *block_counts) {
	const long long base = (long long)blockIdx.x * kTile;
	uint32_t c = 0;
	for (int i = threadIdx.x; i < kTile; i += kThreads)
		if (base + i < n) c += selected(px[base + i], mask_mode, min_bright);
	__shared__ uint32_t s[kThreads / 32];
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < kThreads / 32; ++w) t += s[w];
		block_counts[blockIdx.x] = t;
	}
}

// exclusive scan of `nblocks` u64 counts in place by one 1024-thread block (tiles of 1024 + carry)
__global__ void __launch_bounds__(1024) select_scan_kernel(unsigned long long *block_counts, long long nblocks,
                                                           unsigned long long *total) {
	__shared__ unsigned long long s[1024];
	__shared__ unsigned long long carry;
	if (threadIdx.x == 0) carry = 0ull;
	__syncthreads();
	for (long long t0 = 0; t0 < nblocks; t0 += 1024) {
		const long long i = t0 + threadIdx.x;
		const unsigned long long v = i < nblocks ? block_counts[i] : 0ull;
		s[threadIdx.x] = v;
		__syncthreads();
		for (int o = 1; o < 1024; o <<= 1) {
			const unsigned long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0ull;
			__syncthreads();
			s[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < nblocks) block_counts[i] = carry + s[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry += s[1023];
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kThreads) select_scatter_kernel(
    const uint32_t *__restrict__ px, long long n, int mask_mode, int min_bright,
    const unsigned long long *__restrict__ block_offsets, uint32_t *__restrict__ out_px,
    long long *__restrict__ out_index, long long capacity) {
	const long long base = (long long)blockIdx.x * kTile;
	__shared__ uint32_t warp_cnt[kThreads / 32];
	__shared__ unsigned long long running;
	if (threadIdx.x == 0) running = block_offsets[blockIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int i0 = 0; i0 < kTile; i0 += kThreads) {
		const long long i = base + i0 + threadIdx.x;
		const uint32_t v = i < n ? px[i] : 0u;
		const bool sel = i < n && selected(v, mask_mode, min_bright);
		const uint32_t m = __ballot_sync(0xffffffffu, sel);
		if (lane == 0) warp_cnt[w] = __popc(m);
		__syncthreads();
		unsigned long long off = running;
		for (int j = 0; j < w; ++j) off += warp_cnt[j];
		if (sel) {
			const long long pos = (long long)off + __popc(m & ((1u << lane) - 1u));
			if (pos < capacity) {
				if (out_px) out_px[pos] = v;
				if (out_index) out_index[pos] = i;
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (int j = 0; j < kThreads / 32; ++j) t += warp_cnt[j];
			running += t;
		}
		__syncthreads();
	}
}

// per-byte histograms of the selected pixels: hist[c*256 + v], c = 0..2
__global__ void __launch_bounds__(kThreads) channel_hist_kernel(const uint32_t *__restrict__ px, long long n,
                                                                int mask_mode, int min_bright,
                                                                unsigned long long *hist) {
	__shared__ uint32_t h[kThreads / 32][768];  // one copy per warp (24 KB): fewer same-bin collisions
	for (int i = threadIdx.x; i < (kThreads / 32) * 768; i += kThreads) (&h[0][0])[i] = 0u;
	__syncthreads();
	uint32_t *mine = h[threadIdx.x >> 5];
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t w = px[i];
		if (!selected(w, mask_mode, min_bright)) continue;
		atomicAdd(mine + (w & 0xFFu), 1u);
		atomicAdd(mine + 256 + ((w >> 8) & 0xFFu), 1u);
		atomicAdd(mine + 512 + ((w >> 16) & 0xFFu), 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 768; i += kThreads) {
		unsigned long long t = 0;
		for (int w = 0; w < kThreads / 32; ++w) t += h[w][i];
		if (t) atomicAdd(hist + i, t);
	}
}

__global__ void __launch_bounds__(kThreads) gather_kernel(const uint32_t *__restrict__ px, long long n,
                                                          const long long *__restrict__ idx, long long m,
                                                          uint32_t *__restrict__ out) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < m; i += stride) {
		const long long j = idx[i];
		out[i] = (j >= 0 && j < n) ? px[j] : 0u;
	}
}

// marks mat[a*256 + b] = 1 for every valid pixel
__global__ void __launch_bounds__(kThreads) cooccurrence_kernel(const uint8_t *__restrict__ la,
                                                                const uint8_t *__restrict__ lb, long long n,
                                                                const uint32_t *__restrict__ selpx, int mask_mode,
                                                                int min_bright, uint32_t *mat) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t a = la[i], b = lb[i];
		if (!label_valid(selpx, i, mask_mode, min_bright, a, 256)) continue;
		uint32_t *p = mat + a * 256u + b;
		if (!__ldcg(p)) *p = 1u;  // benign race: every writer stores the same value
	}
}

} // namespace
} // namespace cs

using namespace cs;

#define CS_STREAM ((cudaStream_t)stream)

extern "C" int cs_select_compact_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, int mask_mode, int min_bright,
                                     uint8_t *d_out_px, int64_t *d_out_index, int64_t capacity,
                                     unsigned long long *d_count, void *stream) {
	CS_REQUIRE(ctx && d_px && d_count, "null pointer");
	CS_REQUIRE(mask_mode == 0 || mask_mode == 1, "mask_mode must be 0 or 1");
	CS_REQUIRE(n >= 0 && capacity >= 0, "n and capacity must be >= 0");
	const long long nblocks = (n + kTile - 1) / kTile;
	CS_REQUIRE(nblocks <= (long long)kMaxPartialBlocks * kMaxPartialVals, "n too large for the scan scratch");
	CS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), CS_STREAM));
	if (n == 0) return 0;
	unsigned long long *blk = reinterpret_cast<unsigned long long *>(ctx->d_partials);
	const uint32_t *px = reinterpret_cast<const uint32_t *>(d_px);
	select_count_kernel<<<(unsigned)nblocks, kThreads, 0, CS_STREAM>>>(px, n, mask_mode, min_bright, blk);
	select_scan_kernel<<<1, 1024, 0, CS_STREAM>>>(blk, nblocks, d_count);
	if (d_out_px || d_out_index)
		select_scatter_kernel<<<(unsigned)nblocks, kThreads, 0, CS_STREAM>>>(
		    px, n, mask_mode, min_bright, blk, reinterpret_cast<uint32_t *>(d_out_px),
		    reinterpret_cast<long long *>(d_out_index), capacity);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_channel_hist_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, int mask_mode, int min_bright,
                                   unsigned long long *d_hist768, void *stream) {
	CS_REQUIRE(ctx && d_px && d_hist768, "null pointer");
	CS_REQUIRE(mask_mode == 0 || mask_mode == 1, "mask_mode must be 0 or 1");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_hist768, 0, 768 * sizeof(unsigned long long), CS_STREAM));
	if (n == 0) return 0;
	channel_hist_kernel<<<grid_for(ctx, (n + kThreads - 1) / kThreads, 4), kThreads, 0, CS_STREAM>>>(
	    reinterpret_cast<const uint32_t *>(d_px), n, mask_mode, min_bright, d_hist768);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_gather_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const int64_t *d_index, int64_t m,
                             uint8_t *d_out_px, void *stream) {
	CS_REQUIRE(ctx && d_px && d_index && d_out_px, "null pointer");
	CS_REQUIRE(n >= 0 && m >= 0, "n and m must be >= 0");
	if (m == 0) return 0;
	gather_kernel<<<grid_for(ctx, (m + kThreads - 1) / kThreads, 8), kThreads, 0, CS_STREAM>>>(
	    reinterpret_cast<const uint32_t *>(d_px), n, reinterpret_cast<const long long *>(d_index), m,
	    reinterpret_cast<uint32_t *>(d_out_px));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_label_cooccurrence_u8(cs_ctx *ctx, const uint8_t *d_labels_a, const uint8_t *d_labels_b,
                                        int64_t n, const uint8_t *d_selpx, int mask_mode, int min_bright,
                                        uint32_t *d_matrix, void *stream) {
	CS_REQUIRE(ctx && d_labels_a && d_labels_b && d_matrix, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_matrix, 0, 256 * 256 * sizeof(uint32_t), CS_STREAM));
	if (n == 0) return 0;
	cooccurrence_kernel<<<grid_for(ctx, (n + kThreads - 1) / kThreads, 8), kThreads, 0, CS_STREAM>>>(
	    d_labels_a, d_labels_b, n, reinterpret_cast<const uint32_t *>(d_selpx), mask_mode, min_bright, d_matrix);
	CS_CUDA(cudaGetLastError());
	return 0;
}
