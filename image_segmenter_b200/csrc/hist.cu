// hist.cu — the integer passes behind median-cut / "octree" / posterize / statistics, sm_100a.
//
// K5 colour histogram + fold + compaction  (Pillow create_pixel_hash, PIL/_imaging Quant.c,
//                                           reached from color_simplify.py:145, 201)
// K6 box sums + nearest-palette map         (compute_palette_from_median_cut,
//                                           map_image_pixels_from_median_box)
// K7 posterize                              (color_simplify.py:255-261, 274)
// K8 statistics / mask counts               (color_simplify.py:44-70, 363-376)
// All integer, bit-exact; HBM-bound streaming reads of RGBA8 with 16-byte loads where the
// pixel count allows, atomics only on L2-resident tables and warp-aggregated first.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t rgb_key(uint32_t w) {  // (r<<16)|(g<<8)|b from RGBA8888 LE
	return ((w & 0xFFu) << 16) | (w & 0xFF00u) | ((w >> 16) & 0xFFu);
}
__device__ __forceinline__ uint32_t cell_key(uint32_t rgbkey, int shift) {
	const int bits = 8 - shift;
	const uint32_t r = (rgbkey >> 16) >> shift, g = ((rgbkey >> 8) & 0xFFu) >> shift, b = (rgbkey & 0xFFu) >> shift;
	return (r << (2 * bits)) | (g << bits) | b;
}

// K5.  Each block owns a CONTIGUOUS chunk of the image (spatial coherence = few colours per chunk) and a
// 2048-entry shared-memory table {colour key, count}: lanes of a warp holding the same colour elect a
// leader (__match_any_sync), the leader claims / finds the colour's slot with one shared CAS and adds the
// warp's count there; only colours whose slot is taken by another colour go to the global table.  The
// table is flushed with one global atomic per occupied slot.  On few-colour images (the application
// re-quantises already simplified images) the 2^24-bin global table sees a handful of atomics per block
// instead of one per warp-colour; on high-entropy images a block whose table mostly misses switches the
// table off and behaves like the plain warp-aggregated kernel.
constexpr int kHistSlots = 2048;
constexpr uint32_t kHistEmpty = 0xFFFFFFFFu;  // not a 24-bit key

__device__ __forceinline__ void hist_add_tab(uint32_t *hist, uint32_t *tkey, uint32_t *tcnt, bool use_tab, uint32_t key,
                                             bool active, uint32_t &hits, uint32_t &misses) {
	const uint32_t act = __ballot_sync(0xffffffffu, active);
	if (!active) return;
	const uint32_t peers = __match_any_sync(act, key);
	if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) != 0u) return;  // not the leader of its colour
	const uint32_t c = (uint32_t)__popc(peers);
	if (use_tab) {
		const uint32_t slot = (key * 2654435761u) >> 21;  // 2048 slots
		const uint32_t old = atomicCAS(tkey + slot, kHistEmpty, key);
		if (old == kHistEmpty || old == key) {
			atomicAdd(tcnt + slot, c);
			++hits;
			return;
		}
		++misses;
	}
	atomicAdd(hist + key, c);
}

__global__ void __launch_bounds__(kThreads) hist_rgb24_kernel(const uint32_t *__restrict__ rgba,
                                                              long long n, uint32_t *hist, int vec_ok) {
	__shared__ uint32_t tkey[kHistSlots], tcnt[kHistSlots];
	__shared__ int s_use_tab;
	for (int i = threadIdx.x; i < kHistSlots; i += kThreads) { tkey[i] = kHistEmpty; tcnt[i] = 0u; }
	if (threadIdx.x == 0) s_use_tab = 1;
	__syncthreads();
	// contiguous chunk of whole 128-pixel groups per block (keeps warps whole and 16-byte loads aligned)
	const long long groups = (n + 127) >> 7;
	const long long per = (groups + gridDim.x - 1) / gridDim.x;
	const long long g0 = (long long)blockIdx.x * per, g1 = g0 + per < groups ? g0 + per : groups;
	uint32_t hits = 0, misses = 0;
	int round = 0;
	for (long long gi = g0 + (threadIdx.x >> 5); gi < g1; gi += kThreads / 32, ++round) {
		const long long i0 = (gi << 7) + ((threadIdx.x & 31) << 2);  // 4 consecutive pixels per lane
		uint32_t w[4] = {0u, 0u, 0u, 0u};
		if (vec_ok && i0 + 3 < n) {
			const uint4 px = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba + i0));
			w[0] = px.x; w[1] = px.y; w[2] = px.z; w[3] = px.w;
		} else {
#pragma unroll
			for (int q = 0; q < 4; ++q)
				if (i0 + q < n) w[q] = rgba[i0 + q];
		}
		const bool use_tab = *reinterpret_cast<volatile int *>(&s_use_tab) != 0;
#pragma unroll
		for (int q = 0; q < 4; ++q) hist_add_tab(hist, tkey, tcnt, use_tab, rgb_key(w[q]), i0 + q < n, hits, misses);
		// after 16 rounds a warp whose leaders mostly miss turns the table off for the block
		if (round == 16) {
			const uint32_t h = __reduce_add_sync(0xffffffffu, hits), m = __reduce_add_sync(0xffffffffu, misses);
			if ((threadIdx.x & 31) == 0 && m > 3u * h) s_use_tab = 0;
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < kHistSlots; i += kThreads)
		if (tcnt[i]) atomicAdd(hist + tkey[i], tcnt[i]);
}

__global__ void __launch_bounds__(kThreads) hist_fold_kernel(const uint32_t *__restrict__ hist,
                                                             int shift, uint32_t *cells) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;  // grid covers 2^24 bins exactly
	const uint32_t c = hist[i];
	if (c) atomicAdd(cells + cell_key(i, shift), c);
}

__global__ void __launch_bounds__(kThreads) count_nonzero_kernel(const uint32_t *__restrict__ v,
                                                                 long long n, uint32_t *out) {
	uint32_t c = 0;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) c += v[i] != 0u;
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// order-preserving compaction of the non-empty bins: chunk counts -> scan -> scatter
constexpr int kChunk = 4096;  // bins per block
__global__ void __launch_bounds__(kThreads) compact_count_kernel(const uint32_t *__restrict__ cells,
                                                                 long long nbins, uint32_t *chunk_counts) {
	const long long base = (long long)blockIdx.x * kChunk;
	uint32_t c = 0;
	for (int i = threadIdx.x; i < kChunk; i += kThreads)
		if (base + i < nbins) c += cells[base + i] != 0u;
	__shared__ uint32_t s[kThreads / 32];
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int w = 0; w < kThreads / 32; ++w) t += s[w];
		chunk_counts[blockIdx.x] = t;
	}
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(uint32_t *chunk_counts, int nchunks, uint32_t *total) {
	// exclusive scan of <= 4096 chunk counts by one block (4 per thread)
	__shared__ uint32_t s[1024];
	uint32_t v[4], sum = 0;
	for (int j = 0; j < 4; ++j) {
		const int i = threadIdx.x * 4 + j;
		v[j] = i < nchunks ? chunk_counts[i] : 0u;
		sum += v[j];
	}
	s[threadIdx.x] = sum;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
		__syncthreads();
		s[threadIdx.x] += t;
		__syncthreads();
	}
	uint32_t run = s[threadIdx.x] - sum;
	for (int j = 0; j < 4; ++j) {
		const int i = threadIdx.x * 4 + j;
		if (i < nchunks) chunk_counts[i] = run;
		run += v[j];
	}
	if (threadIdx.x == 1023) *total = s[1023];
}
__global__ void __launch_bounds__(kThreads) compact_scatter_kernel(const uint32_t *__restrict__ cells,
                                                                   long long nbins,
                                                                   const uint32_t *__restrict__ chunk_offsets,
                                                                   uint32_t capacity, uint32_t *keys,
                                                                   uint32_t *counts) {
	const long long base = (long long)blockIdx.x * kChunk;
	__shared__ uint32_t warp_base[kThreads / 32];
	__shared__ uint32_t running;
	if (threadIdx.x == 0) running = chunk_offsets[blockIdx.x];
	__syncthreads();
	for (int i0 = 0; i0 < kChunk; i0 += kThreads) {
		const long long i = base + i0 + threadIdx.x;
		const uint32_t c = i < nbins ? cells[i] : 0u;
		const uint32_t m = __ballot_sync(0xffffffffu, c != 0u);
		const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
		if (lane == 0) warp_base[w] = __popc(m);
		__syncthreads();
		uint32_t off = running;
		for (int j = 0; j < w; ++j) off += warp_base[j];
		if (c) {
			const uint32_t pos = off + __popc(m & ((1u << lane) - 1u));
			if (pos < capacity) { keys[pos] = (uint32_t)i; counts[pos] = c; }
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (int j = 0; j < kThreads / 32; ++j) t += warp_base[j];
			running += t;
		}
		__syncthreads();
	}
}

// K6a: per-box sums of the unscaled r,g,b and pixel counts from the 2^24-bin histogram
__global__ void __launch_bounds__(kThreads) box_sums_kernel(const uint32_t *__restrict__ hist,
                                                            const uint16_t *__restrict__ cell_box, int shift,
                                                            int n_boxes, unsigned long long *box_acc) {
	__shared__ unsigned long long acc[CS_MAX_K * 4];
	for (int i = threadIdx.x; i < n_boxes * 4; i += kThreads) acc[i] = 0ull;
	__syncthreads();
	const uint32_t stride = gridDim.x * kThreads;
	for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < (1u << 24); i += stride) {
		const uint32_t c = hist[i];
		if (!c) continue;
		const uint32_t box = cell_box[cell_key(i, shift)];
		if (box >= (uint32_t)n_boxes) continue;
		atomicAdd(&acc[box * 4 + 0], (unsigned long long)c * (i >> 16));
		atomicAdd(&acc[box * 4 + 1], (unsigned long long)c * ((i >> 8) & 0xFFu));
		atomicAdd(&acc[box * 4 + 2], (unsigned long long)c * (i & 0xFFu));
		atomicAdd(&acc[box * 4 + 3], (unsigned long long)c);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < n_boxes * 4; i += kThreads)
		if (acc[i]) atomicAdd(box_acc + i, acc[i]);
}

// K6b: pixel -> palette index, Pillow's map_image_pixels_from_median_box rule
__global__ void __launch_bounds__(kThreads) palette_map_kernel(
    const uint32_t *__restrict__ rgba, long long n, const uint16_t *__restrict__ cell_box, int shift,
    const uint8_t *__restrict__ palette, int n_pal, int preserve_alpha, uint32_t *__restrict__ out,
    uint8_t *__restrict__ index) {
	__shared__ int pr[CS_MAX_K], pg[CS_MAX_K], pb[CS_MAX_K];
	for (int i = threadIdx.x; i < n_pal; i += kThreads) {
		pr[i] = palette[3 * i]; pg[i] = palette[3 * i + 1]; pb[i] = palette[3 * i + 2];
	}
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t w = rgba[i];
		const int r = w & 0xFF, g = (w >> 8) & 0xFF, b = (w >> 16) & 0xFF;
		const uint32_t a = w >> 24;
		int own = cell_box[cell_key(rgb_key(w), shift)];
		if (own >= n_pal) own = 0;  // cannot happen for a LUT built from this image
		const int dor = r - pr[own], dog = g - pg[own], dob = b - pb[own];
		const uint32_t d_own = (uint32_t)(dor * dor + dog * dog + dob * dob);
		uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;  // (d << 8) | j, d <= 195075 < 2^18
		for (int j = 0; j < n_pal; ++j) {
			const int dr = r - pr[j], dg = g - pg[j], db = b - pb[j];
			const uint32_t key = ((uint32_t)(dr * dr + dg * dg + db * db) << 8) | (uint32_t)j;
			second = min(second, max(best, key));
			best = min(best, key);
		}
		const uint32_t dmin = best >> 8;
		int pick;
		if (d_own == dmin) {
			pick = own;  // the scan starts at the own box and only a strictly smaller distance replaces it
		} else if ((second >> 8) != dmin) {
			pick = (int)(best & 0xFFu);  // unique minimiser
		} else {
			// several entries tie: Pillow visits them in ascending (palette distance from own, index)
			uint32_t bk = 0xFFFFFFFFu;
			for (int j = 0; j < n_pal; ++j) {
				const int dr = r - pr[j], dg = g - pg[j], db = b - pb[j];
				if ((uint32_t)(dr * dr + dg * dg + db * db) != dmin) continue;
				const int er = pr[own] - pr[j], eg = pg[own] - pg[j], eb = pb[own] - pb[j];
				bk = min(bk, ((uint32_t)(er * er + eg * eg + eb * eb) << 8) | (uint32_t)j);
			}
			pick = (int)(bk & 0xFFu);
		}
		const uint32_t a_out = preserve_alpha ? a : (a > 128u ? 255u : 0u);
		out[i] = (uint32_t)pr[pick] | ((uint32_t)pg[pick] << 8) | ((uint32_t)pb[pick] << 16) | (a_out << 24);
		if (index) index[i] = (uint8_t)pick;
	}
}

// test-then-set of one bit in a presence bitmap (the atomic is skipped once the bit is set)
__device__ __forceinline__ void bitmap_set(uint32_t *bm, uint32_t key) {
	uint32_t *wp = bm + (key >> 5);
	const uint32_t bit = 1u << (key & 31u);
	if (!(__ldcg(wp) & bit)) atomicOr(wp, bit);
}

// K7: 4 pixels per thread (one 16-byte streaming load / store).  The distinct quantised colours are few
// (at most ceil(256/step)^3), so a per-block direct-mapped cache of recently marked keys sits in front of
// the global bitmap: a pixel touches L2 only when its colour is not the one last seen in its cache line.
__device__ __forceinline__ uint32_t posterize_word(const uint32_t *q, uint32_t w, int preserve_alpha, uint32_t *present,
                                                   uint32_t *seen) {
	const uint32_t r = q[w & 0xFFu], g = q[(w >> 8) & 0xFFu], b = q[(w >> 16) & 0xFFu], a = w >> 24;
	const uint32_t a_out = preserve_alpha ? a : (a > 128u ? 255u : 0u);
	if (present) {
		const uint32_t key = (r << 16) | (g << 8) | b;
		const uint32_t slot = (key * 2654435761u) >> 22;  // 1024 entries
		if (seen[slot] != key) {  // benign race: a stale entry only costs one more global test
			seen[slot] = key;
			bitmap_set(present, key);
		}
	}
	return r | (g << 8) | (b << 16) | (a_out << 24);
}

__global__ void __launch_bounds__(kThreads) posterize_kernel(const uint32_t *__restrict__ rgba,
                                                             long long n, int step, int preserve_alpha,
                                                             uint32_t *__restrict__ out, uint32_t *present, int vec_ok) {
	__shared__ uint32_t q[256];
	__shared__ uint32_t seen[1024];
	for (int i = threadIdx.x; i < 256; i += kThreads) q[i] = step > 0 ? (uint32_t)((i / step) * step) & 0xFFu : 0u;
	for (int i = threadIdx.x; i < 1024; i += kThreads) seen[i] = 0xFFFFFFFFu;  // not an RGB key
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	const long long n4 = vec_ok ? (n >> 2) : 0;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
		uint4 px = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i);
		px.x = posterize_word(q, px.x, preserve_alpha, present, seen);
		px.y = posterize_word(q, px.y, preserve_alpha, present, seen);
		px.z = posterize_word(q, px.z, preserve_alpha, present, seen);
		px.w = posterize_word(q, px.w, preserve_alpha, present, seen);
		stg_stream_u4(reinterpret_cast<uint4 *>(out) + i, px);
	}
	for (long long i = (n4 << 2) + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
		out[i] = posterize_word(q, rgba[i], preserve_alpha, present, seen);
}

// K8
__global__ void __launch_bounds__(kThreads) stats_kernel(const uint32_t *__restrict__ rgba, long long n,
                                                         uint32_t *bitmap, unsigned long long *acc, int vec_ok) {
	// per-thread partial moments in 32 bits (<= 2^17 pixels per thread between flushes: 255^2 * 2^15 < 2^32)
	unsigned long long v[7] = {0, 0, 0, 0, 0, 0, 0};
	uint32_t t[7] = {0, 0, 0, 0, 0, 0, 0};
	int pending = 0;
	auto one = [&](uint32_t w) {
		if (bitmap) bitmap_set(bitmap, w);
		if (w >> 24) {
			const uint32_t r = w & 0xFFu, g = (w >> 8) & 0xFFu, b = (w >> 16) & 0xFFu;
			t[0] += 1; t[1] += r; t[2] += g; t[3] += b; t[4] += r * r; t[5] += g * g; t[6] += b * b;
		}
	};
	auto flush = [&]() {
#pragma unroll
		for (int j = 0; j < 7; ++j) { v[j] += t[j]; t[j] = 0; }
		pending = 0;
	};
	const long long stride = (long long)gridDim.x * kThreads;
	const long long n4 = vec_ok ? (n >> 2) : 0;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += 2 * stride) {
		const bool two = i + stride < n4;  // two 16-byte loads in flight per thread
		const uint4 px = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i);
		const uint4 py = two ? ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i + stride) : make_uint4(0u, 0u, 0u, 0u);
		one(px.x); one(px.y); one(px.z); one(px.w);
		if (two) { one(py.x); one(py.y); one(py.z); one(py.w); }
		if (++pending == 4096) flush();
	}
	for (long long i = (n4 << 2) + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		one(rgba[i]);
		if (++pending == 4096) flush();
	}
	flush();
	__shared__ unsigned long long s[7];
	if (threadIdx.x < 7) s[threadIdx.x] = 0ull;
	__syncthreads();
	for (int j = 0; j < 7; ++j) {
		unsigned long long t = v[j];
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if ((threadIdx.x & 31) == 0 && t) atomicAdd(&s[j], t);
	}
	__syncthreads();
	if (threadIdx.x < 7 && s[threadIdx.x]) atomicAdd(acc + threadIdx.x, s[threadIdx.x]);
}

__global__ void __launch_bounds__(kThreads) popcount_kernel(const uint32_t *__restrict__ bm, long long n_words,
                                                            unsigned long long *out) {
	unsigned long long c = 0;
	const long long stride = (long long)gridDim.x * kThreads;
	const long long n4 = n_words >> 2;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
		const uint4 w = reinterpret_cast<const uint4 *>(bm)[i];
		c += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
	}
	if (blockIdx.x == 0 && threadIdx.x < (n_words & 3)) c += __popc(bm[(n4 << 2) + threadIdx.x]);
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// MODE 0: RGBA pixels, brightness = r+g+b (thresholds 90 / 30); MODE 1: HSVA pixels, brightness = v
template <int MODE>
__global__ void __launch_bounds__(kThreads) mask_stats_kernel(const uint32_t *__restrict__ px, long long n,
                                                              int min_bright, int t_hi, int t_lo,
                                                              uint32_t *present, unsigned long long *acc, int vec_ok) {
	unsigned long long c0 = 0, c1 = 0, c2 = 0;
	auto one = [&](uint32_t w) {
		if (!(w >> 24)) return;
		const int br = MODE == 0 ? (int)((w & 0xFFu) + ((w >> 8) & 0xFFu) + ((w >> 16) & 0xFFu)) : (int)((w >> 16) & 0xFFu);
		c0 += 1; c1 += br > t_hi; c2 += br > t_lo;
		if (present && br > min_bright) bitmap_set(present, rgb_key(w));
	};
	const long long stride = (long long)gridDim.x * kThreads;
	const long long n4 = vec_ok ? (n >> 2) : 0;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += 2 * stride) {
		const bool two = i + stride < n4;  // two 16-byte loads in flight per thread
		const uint4 q = ldg_stream_u4(reinterpret_cast<const uint4 *>(px) + i);
		const uint4 r = two ? ldg_stream_u4(reinterpret_cast<const uint4 *>(px) + i + stride) : make_uint4(0u, 0u, 0u, 0u);
		one(q.x); one(q.y); one(q.z); one(q.w);
		if (two) { one(r.x); one(r.y); one(r.z); one(r.w); }
	}
	for (long long i = (n4 << 2) + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) one(px[i]);
	unsigned long long v[3] = {c0, c1, c2};
	for (int j = 0; j < 3; ++j) {
		unsigned long long t = v[j];
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if ((threadIdx.x & 31) == 0 && t) atomicAdd(acc + j, t);
	}
}

// per-label sums of r,g,b and counts (cluster centres "in RGB space", color_simplify.py:996-1000)
__global__ void __launch_bounds__(kThreads) sum_by_label_kernel(const uint32_t *__restrict__ rgba,
                                                                const uint8_t *__restrict__ labels, long long n,
                                                                const uint32_t *__restrict__ selpx, int mask_mode,
                                                                int min_bright, int K, unsigned long long *acc_g) {
	__shared__ unsigned long long acc[CS_MAX_K * 4];
	for (int i = threadIdx.x; i < K * 4; i += kThreads) acc[i] = 0ull;
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t l = labels[i];
		if (l >= (uint32_t)K || !label_valid(selpx, i, mask_mode, min_bright, l, K)) continue;
		const uint32_t w = rgba[i];
		atomicAdd(&acc[l * 4 + 0], (unsigned long long)(w & 0xFFu));
		atomicAdd(&acc[l * 4 + 1], (unsigned long long)((w >> 8) & 0xFFu));
		atomicAdd(&acc[l * 4 + 2], (unsigned long long)((w >> 16) & 0xFFu));
		atomicAdd(&acc[l * 4 + 3], 1ull);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < K * 4; i += kThreads)
		if (acc[i]) atomicAdd(acc_g + i, acc[i]);
}

__global__ void __launch_bounds__(kThreads) merge_labels_kernel(const uint8_t *__restrict__ a,
                                                                const uint8_t *__restrict__ b, long long n,
                                                                const uint32_t *__restrict__ selpx, int mask_mode,
                                                                int min_bright, uint8_t *__restrict__ out) {
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint8_t v = a[i];
		const bool keep = selpx ? px_selected(selpx[i], mask_mode, min_bright) : v != 255;
		out[i] = keep ? v : b[i];
	}
}

} // namespace
} // namespace cs

using namespace cs;

#define CS_STREAM ((cudaStream_t)stream)
#define CS_GRID(n) grid_for(ctx, ((n) + kThreads - 1) / kThreads, 8)

extern "C" int cs_hist_rgb24(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint32_t *d_hist, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_hist, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	if (n == 0) return 0;
	hist_rgb24_kernel<<<grid_for(ctx, ((n + 127) / 128 + 31) / 32, 8), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_rgba), n, d_hist,
	                                                                       (int)(((uintptr_t)d_rgba & 15u) == 0));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_hist_fold(cs_ctx *ctx, const uint32_t *d_hist, int shift, uint32_t *d_cells,
                            uint32_t *d_ncells, void *stream) {
	CS_REQUIRE(ctx && d_hist && d_cells && d_ncells, "null pointer");
	CS_REQUIRE(shift >= 0 && shift <= 7, "shift must be in [0,7]");
	const long long ncell = 1LL << (3 * (8 - shift));
	CS_CUDA(cudaMemsetAsync(d_cells, 0, ncell * sizeof(uint32_t), CS_STREAM));
	CS_CUDA(cudaMemsetAsync(d_ncells, 0, sizeof(uint32_t), CS_STREAM));
	hist_fold_kernel<<<(1 << 24) / kThreads, kThreads, 0, CS_STREAM>>>(d_hist, shift, d_cells);
	count_nonzero_kernel<<<CS_GRID(ncell), kThreads, 0, CS_STREAM>>>(d_cells, ncell, d_ncells);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_hist_compact(cs_ctx *ctx, const uint32_t *d_cells, int64_t nbins, uint32_t *d_keys,
                               uint32_t *d_counts, uint32_t capacity, uint32_t *d_n, void *stream) {
	CS_REQUIRE(ctx && d_cells && d_keys && d_counts && d_n, "null pointer");
	CS_REQUIRE(nbins > 0 && nbins <= (1LL << 24), "nbins must be in (0, 2^24]");
	const int nchunks = (int)((nbins + kChunk - 1) / kChunk);  // <= 4096
	uint32_t *chunk = reinterpret_cast<uint32_t *>(ctx->d_partials);  // reuse the partials scratch
	compact_count_kernel<<<nchunks, kThreads, 0, CS_STREAM>>>(d_cells, nbins, chunk);
	compact_scan_kernel<<<1, 1024, 0, CS_STREAM>>>(chunk, nchunks, d_n);
	compact_scatter_kernel<<<nchunks, kThreads, 0, CS_STREAM>>>(d_cells, nbins, chunk, capacity, d_keys, d_counts);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_box_sums(cs_ctx *ctx, const uint32_t *d_hist, const uint16_t *d_cell_box, int shift,
                           int n_boxes, unsigned long long *d_box_acc, void *stream) {
	CS_REQUIRE(ctx && d_hist && d_cell_box && d_box_acc, "null pointer");
	CS_REQUIRE(shift >= 0 && shift <= 7, "shift must be in [0,7]");
	CS_REQUIRE(n_boxes >= 1 && n_boxes <= CS_MAX_K, "n_boxes must be in [1,256]");
	CS_CUDA(cudaMemsetAsync(d_box_acc, 0, sizeof(unsigned long long) * 4 * n_boxes, CS_STREAM));
	box_sums_kernel<<<ctx->sm_count * 4, kThreads, 0, CS_STREAM>>>(d_hist, d_cell_box, shift, n_boxes, d_box_acc);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_palette_map_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n,
                                    const uint16_t *d_cell_box, int shift, const uint8_t *d_palette_rgb,
                                    int n_pal, int preserve_alpha, uint8_t *d_rgba_out, uint8_t *d_index,
                                    void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_cell_box && d_palette_rgb && d_rgba_out, "null pointer");
	CS_REQUIRE(shift >= 0 && shift <= 7, "shift must be in [0,7]");
	CS_REQUIRE(n_pal >= 1 && n_pal <= CS_MAX_K, "n_pal must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	if (n == 0) return 0;
	palette_map_kernel<<<CS_GRID(n), kThreads, 0, CS_STREAM>>>(
	    reinterpret_cast<const uint32_t *>(d_rgba), n, d_cell_box, shift, d_palette_rgb, n_pal, preserve_alpha,
	    reinterpret_cast<uint32_t *>(d_rgba_out), d_index);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_posterize_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int step, int preserve_alpha,
                                  uint8_t *d_rgba_out, uint32_t *d_present, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_rgba_out, "null pointer");
	CS_REQUIRE(n >= 0 && step >= 0 && step <= 256, "bad n or step");
	if (n == 0) return 0;
	const int vec_ok = (((uintptr_t)d_rgba | (uintptr_t)d_rgba_out) & 15u) == 0;
	posterize_kernel<<<CS_GRID(n / 4 + 1), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_rgba), n, step,
	                                                                preserve_alpha, reinterpret_cast<uint32_t *>(d_rgba_out),
	                                                                d_present, vec_ok);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_stats_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint32_t *d_bitmap,
                              unsigned long long *d_acc, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_acc, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_acc, 0, 8 * sizeof(unsigned long long), CS_STREAM));
	if (n == 0) return 0;
	stats_kernel<<<CS_GRID(n / 4 + 1), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_rgba), n, d_bitmap, d_acc,
	                                                            (int)(((uintptr_t)d_rgba & 15u) == 0));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_bitmap_popcount(cs_ctx *ctx, const uint32_t *d_bitmap, int64_t n_words,
                                  unsigned long long *d_count, void *stream) {
	CS_REQUIRE(ctx && d_bitmap && d_count, "null pointer");
	CS_REQUIRE(n_words >= 0, "n_words must be >= 0");
	CS_REQUIRE(((uintptr_t)d_bitmap & 15u) == 0, "bitmap must be 16-byte aligned");
	CS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), CS_STREAM));
	if (n_words == 0) return 0;
	popcount_kernel<<<CS_GRID(n_words / 4 + 1), kThreads, 0, CS_STREAM>>>(d_bitmap, n_words, d_count);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_mask_stats_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                                   uint32_t *d_present, unsigned long long *d_acc, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_acc, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_acc, 0, 4 * sizeof(unsigned long long), CS_STREAM));
	if (n == 0) return 0;
	mask_stats_kernel<0><<<CS_GRID(n / 4 + 1), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_rgba), n,
	                                                                    min_rgb_sum, 90, 30, d_present, d_acc,
	                                                                    (int)(((uintptr_t)d_rgba & 15u) == 0));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_mask_stats_hsv8(cs_ctx *ctx, const uint8_t *d_hsva, int64_t n, int min_v,
                                  uint32_t *d_present, unsigned long long *d_acc, void *stream) {
	CS_REQUIRE(ctx && d_hsva && d_acc, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_acc, 0, 4 * sizeof(unsigned long long), CS_STREAM));
	if (n == 0) return 0;
	mask_stats_kernel<1><<<CS_GRID(n / 4 + 1), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_hsva), n,
	                                                                    min_v, 30, 10, d_present, d_acc,
	                                                                    (int)(((uintptr_t)d_hsva & 15u) == 0));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_sum_by_label_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, const uint8_t *d_labels, int64_t n,
                                     const uint8_t *d_selpx, int mask_mode, int min_bright, int K,
                                     unsigned long long *d_acc, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_labels && d_acc, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0, "bad K or n");
	CS_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(unsigned long long) * 4 * K, CS_STREAM));
	if (n == 0) return 0;
	sum_by_label_kernel<<<grid_for(ctx, (n + kThreads - 1) / kThreads, 4), kThreads, 0, CS_STREAM>>>(
	    reinterpret_cast<const uint32_t *>(d_rgba), d_labels, n, reinterpret_cast<const uint32_t *>(d_selpx), mask_mode,
	    min_bright, K, d_acc);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_merge_labels_u8(cs_ctx *ctx, const uint8_t *d_primary, const uint8_t *d_fallback, int64_t n,
                                  const uint8_t *d_selpx, int mask_mode, int min_bright, uint8_t *d_out,
                                  void *stream) {
	CS_REQUIRE(ctx && d_primary && d_fallback && d_out, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	if (n == 0) return 0;
	merge_labels_kernel<<<CS_GRID(n), kThreads, 0, CS_STREAM>>>(d_primary, d_fallback, n,
	                                                          reinterpret_cast<const uint32_t *>(d_selpx), mask_mode,
	                                                          min_bright, d_out);
	CS_CUDA(cudaGetLastError());
	return 0;
}
