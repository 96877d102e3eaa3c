// kmeanspp.cu — k-means++ seeding passes (sklearn/cluster/_kmeans.py:180-278 `_kmeans_plusplus`),
// sm_100a.  KMeans.fit seeds every one of its n_init runs this way before the Lloyd loop the
// reference reaches at app/processing/color_simplify.py:79-80, 992-993.
//
// sklearn's round, per new centre:  rand_vals = uniform(T) * pot;  candidates =
// searchsorted(cumsum(closest_dist_sq), rand_vals);  distance of every sample to each candidate;
// pot_t = sum_i min(closest_i, d_it);  keep the candidate with the smallest pot.  The random stream
// (RandomState(42)) and the tiny decisions stay on the host; the three O(N) steps are kernels over the
// compacted packed pixels (cs_select_compact_px8 order = the reference's row order):
//   kpp_eval    pot_t for T candidates in one pass (per-block partial sums, summed by the host in order)
//   kpp_update  closest_i = min(closest_i, d_i(winner)) and the per-tile sums of the new closest array
//   kpp_locate  the sequential-order crossing index inside one tile (host picks the tile from the
//               cumulative tile sums): numpy's searchsorted(cumsum(.), v) without an O(N) serial scan.
// Features are fp64 functions of one byte (3 x 256 table; identity for RGB, the weighted HSV features
// of color_simplify.py:969-981 exactly as float64), distances are the direct fp64 formula.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;
constexpr int kTile = 4096;
constexpr int kMaxTrials = 8;  // 2 + int(ln K) <= 7 for K <= 256

__device__ __forceinline__ void feat_of(const double *lut, uint32_t w, double &x, double &y, double &z) {
	x = lut[w & 0xFFu]; y = lut[256 + ((w >> 8) & 0xFFu)]; z = lut[512 + ((w >> 16) & 0xFFu)];
}
__device__ __forceinline__ double dist2(double x, double y, double z, const double *c) {
	const double dx = x - c[0], dy = y - c[1], dz = z - c[2];
	return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---- the three passes, for SEVERAL initialisations at once (gridDim.y = initialisation; n_batch = 1 is the plain case) ----
// KMeans.fit seeds n_init runs (sklearn/cluster/_kmeans.py:1506-1514); their random draws are fixed in count, so
// the host advances all of them in lockstep (engine.Engine.kmeanspp_seeds) and every pass is ONE launch.
// Per-initialisation arrays are stacked: closest [B][n], tile_sums [B][ntiles], block_pots [B][pot_stride][8],
// candidates [B][8][3], tile / index [B][8], prefix and val [B][16] (prefix in 0..7, val in 8..15).
__global__ void __launch_bounds__(kThreads) kpp_eval_batched_kernel(const uint32_t *__restrict__ px, long long n,
                                                                    const double *__restrict__ lut_g,
                                                                    const double *__restrict__ rows,
                                                                    const double *__restrict__ cands, int n_cand,
                                                                    const double *__restrict__ closest_all,
                                                                    double *__restrict__ block_pots_all, int pot_stride) {
	__shared__ double lut[768];
	__shared__ double red[kThreads / 32][kMaxTrials];
	__shared__ double cf[kMaxTrials][3];
	const int b = blockIdx.y;
	if (!rows)
		for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = lut_g[i];
	if (threadIdx.x < kMaxTrials * 3) cf[threadIdx.x / 3][threadIdx.x % 3] = cands[(size_t)b * kMaxTrials * 3 + threadIdx.x];
	__syncthreads();
	const double *closest = closest_all + (size_t)b * n;
	double acc[kMaxTrials];
#pragma unroll
	for (int t = 0; t < kMaxTrials; ++t) acc[t] = 0.0;
	const long long per = (n + gridDim.x - 1) / gridDim.x;
	const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
	for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
		double x, y, z;
		if (rows) { x = rows[3 * i]; y = rows[3 * i + 1]; z = rows[3 * i + 2]; }  // fp64 feature rows (rows64.cu callers)
		else feat_of(lut, px[i], x, y, z);
		const double c0 = closest[i];
#pragma unroll
		for (int t = 0; t < kMaxTrials; ++t)
			if (t < n_cand) acc[t] += fmin(c0, dist2(x, y, z, cf[t]));
	}
#pragma unroll
	for (int t = 0; t < kMaxTrials; ++t) {
		double v = acc[t];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
	}
	__syncthreads();
	if (threadIdx.x < kMaxTrials) {
		double v = 0.0;
		for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
		block_pots_all[((size_t)b * pot_stride + blockIdx.x) * kMaxTrials + threadIdx.x] = v;
	}
}

// centre of initialisation b = candidate pick[b] of cands[b] (pick NULL: candidate 0)
__global__ void __launch_bounds__(kThreads) kpp_update_batched_kernel(const uint32_t *__restrict__ px, long long n,
                                                                      const double *__restrict__ lut_g,
                                                                      const double *__restrict__ rows,
                                                                      const double *__restrict__ cands,
                                                                      const int *__restrict__ pick, int first,
                                                                      double *__restrict__ closest_all,
                                                                      double *__restrict__ tile_sums_all) {
	__shared__ double lut[768];
	__shared__ double red[kThreads / 32];
	const int b = blockIdx.y;
	if (!rows)
		for (int i = threadIdx.x; i < 768; i += kThreads) lut[i] = lut_g[i];
	__syncthreads();
	const double *cp = cands + ((size_t)b * kMaxTrials + (pick ? pick[b] : 0)) * 3;
	const double c[3] = {cp[0], cp[1], cp[2]};
	double *closest = closest_all + (size_t)b * n;
	const long long base = (long long)blockIdx.x * kTile;
	double acc = 0.0;
	for (int j = threadIdx.x; j < kTile; j += kThreads) {
		const long long i = base + j;
		if (i >= n) break;
		double x, y, z;
		if (rows) { x = rows[3 * i]; y = rows[3 * i + 1]; z = rows[3 * i + 2]; }
		else feat_of(lut, px[i], x, y, z);
		double d = dist2(x, y, z, c);
		if (!first) d = fmin(closest[i], d);
		closest[i] = d;
		acc += d;
	}
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double v = 0.0;
		for (int w = 0; w < kThreads / 32; ++w) v += red[w];
		tile_sums_all[(size_t)b * gridDim.x + blockIdx.x] = v;
	}
}

// one block per initialisation, one thread per query; the walk loads 8 values ahead of the dependent adds
__global__ void __launch_bounds__(32) kpp_locate_batched_kernel(const double *__restrict__ closest_all, long long n,
                                                                const long long *__restrict__ tile,
                                                                const double *__restrict__ prefix_val, int nq,
                                                                const uint32_t *__restrict__ px, long long *__restrict__ out,
                                                                uint32_t *__restrict__ out_px) {
	const int b = blockIdx.x, q = threadIdx.x;
	if (q >= nq) return;
	const double *closest = closest_all + (size_t)b * n;
	const long long lo = tile[b * kMaxTrials + q] * (long long)kTile;
	const long long hi = lo + kTile < n ? lo + kTile : n;
	double run = prefix_val[b * 2 * kMaxTrials + q];
	const double val = prefix_val[b * 2 * kMaxTrials + kMaxTrials + q];
	long long found = hi - 1;
	bool done = false;
	for (long long i = lo; i < hi && !done; i += 8) {
		double v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) v[j] = i + j < hi ? closest[i + j] : 0.0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (!done && i + j < hi) {
				run = __dadd_rn(run, v[j]);
				if (run >= val) { found = i + j; done = true; }
			}
		}
	}
	out[b * kMaxTrials + q] = found;
	if (out_px) out_px[b * kMaxTrials + q] = px[found];
}

// potentials of the candidates = per-block partials added in block order (as the host did with
// block_pots.sum(axis=0)); pick = first minimum (np.argmin)
__global__ void __launch_bounds__(32) kpp_pick_batched_kernel(const double *__restrict__ block_pots_all, int pot_stride, int nb,
                                                              int n_cand, int *__restrict__ pick, double *__restrict__ pot) {
	__shared__ double pots[kMaxTrials];
	const int b = blockIdx.x, t = threadIdx.x;
	if (t < n_cand) {
		double s = 0.0;
		for (int r = 0; r < nb; ++r) s = __dadd_rn(s, block_pots_all[((size_t)b * pot_stride + r) * kMaxTrials + t]);
		pots[t] = s;
	}
	__syncwarp();
	if (t == 0) {
		int best = 0;
		for (int c = 1; c < n_cand; ++c)
			if (pots[c] < pots[best]) best = c;
		pick[b] = best;
		pot[b] = pots[best];
	}
}

// One whole "draw" of a round on the device (what the host did between two passes): rand_vals = uniform * pot,
// the tile of each value from the sequential cumulative sum of the tile sums (np.cumsum + np.searchsorted(...,
// side="left"), clipped to the last tile), the crossing index inside that tile (kpp_locate's walk), and the
// candidate's features.  One block per initialisation, one WARP per trial: the lanes load 32 consecutive values at
// a time (coalesced, two chunks ahead) and every lane adds them in index order through shuffles — the same
// running sums in the same order as np.cumsum, so the result equals the host's bit for bit, without a
// load latency per element.
__device__ __forceinline__ long long kpp_ordered_cross(const double *__restrict__ a, long long lo, long long hi, double start,
                                                       double val, double &before, int lane) {
	// first i in [lo, hi) with start + a[lo] + ... + a[i] >= val (sums in index order); hi if none.  `before` = the
	// running sum in front of that element (or of everything, if none)
	double run = start;
	double nxt0 = lo + lane < hi ? a[lo + lane] : 0.0;
	double nxt1 = lo + 32 + lane < hi ? a[lo + 32 + lane] : 0.0;
	for (long long base = lo; base < hi; base += 32) {
		const double cur = nxt0;
		nxt0 = nxt1;
		nxt1 = base + 64 + lane < hi ? a[base + 64 + lane] : 0.0;
		const int cnt = hi - base < 32 ? (int)(hi - base) : 32;
		for (int j = 0; j < cnt; ++j) {
			const double v = __shfl_sync(0xffffffffu, cur, j);
			const double nx = __dadd_rn(run, v);
			if (nx >= val) { before = run; return base + j; }
			run = nx;
		}
	}
	before = run;
	return hi;
}

__global__ void __launch_bounds__(32 * kMaxTrials) kpp_draw_batched_kernel(const double *__restrict__ closest_all, long long n,
                                                              const double *__restrict__ tile_sums_all, long long ntiles,
                                                              const double *__restrict__ pot, const double *__restrict__ uniforms,
                                                              int nq, const uint32_t *__restrict__ px, const double *__restrict__ lut,
                                                              const double *__restrict__ rows, long long *__restrict__ cand_idx,
                                                              double *__restrict__ cands) {
	const int b = blockIdx.x, q = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (q >= nq) return;
	const double *ts = tile_sums_all + (size_t)b * ntiles;
	const double val = __dmul_rn(uniforms[b * kMaxTrials + q], pot[b]);
	double prefix;
	long long tile = kpp_ordered_cross(ts, 0, ntiles, 0.0, val, prefix, lane);
	if (tile >= ntiles) {  // value beyond the last cumulative sum: the last tile, entered with the sum before it
		tile = ntiles - 1;
		double dummy;
		kpp_ordered_cross(ts, 0, ntiles - 1, 0.0, 1e300, dummy, lane);
		prefix = dummy;
	}
	const double *closest = closest_all + (size_t)b * n;
	const long long lo = tile * (long long)kTile;
	const long long hi = lo + kTile < n ? lo + kTile : n;
	double before;
	long long found = kpp_ordered_cross(closest, lo, hi, prefix, val, before, lane);
	if (found >= hi) found = hi - 1;
	if (lane == 0) {
		cand_idx[b * kMaxTrials + q] = found;
		double x, y, z;
		if (rows) { x = rows[3 * found]; y = rows[3 * found + 1]; z = rows[3 * found + 2]; }
		else feat_of(lut, px[found], x, y, z);
		double *c = cands + ((size_t)b * kMaxTrials + q) * 3;
		c[0] = x; c[1] = y; c[2] = z;
	}
}

// the round's winner into the per-initialisation result arrays: index [B][K], centre [B][K][3]
__global__ void kpp_record_batched_kernel(const long long *__restrict__ cand_idx, const double *__restrict__ cands,
                                          const int *__restrict__ pick, int slot, int K, int n_batch,
                                          long long *__restrict__ idx_out, double *__restrict__ cent_out) {
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= n_batch) return;
	const int w = pick[b];
	idx_out[(size_t)b * K + slot] = cand_idx[b * kMaxTrials + w];
	for (int j = 0; j < 3; ++j) cent_out[((size_t)b * K + slot) * 3 + j] = cands[((size_t)b * kMaxTrials + w) * 3 + j];
}

} // namespace
} // namespace cs

using namespace cs;

#define CS_STREAM ((cudaStream_t)stream)

// ---- entry points: n_batch initialisations per launch (see the kernels above for the layouts) ----
extern "C" int cs_kpp_eval_batched(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const double *d_lut768, const double *d_rows,
                                   const double *d_cands, int n_cand, const double *d_closest, double *d_block_pots,
                                   int pot_stride, int n_batch, int *h_n_blocks, void *stream) {
	CS_REQUIRE(ctx && (d_rows || (d_px && d_lut768)) && d_cands && d_closest && d_block_pots && h_n_blocks, "null pointer");
	CS_REQUIRE(n > 0 && n_cand >= 1 && n_cand <= kMaxTrials && n_batch >= 1 && n_batch <= 65535, "bad n, n_cand or n_batch");
	int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 4);
	if (grid > pot_stride) grid = pot_stride;
	CS_REQUIRE(grid >= 1, "pot_stride must be >= 1");
	*h_n_blocks = grid;
	kpp_eval_batched_kernel<<<dim3(grid, n_batch), kThreads, 0, CS_STREAM>>>(reinterpret_cast<const uint32_t *>(d_px), n, d_lut768,
	                                                                         d_rows, d_cands, n_cand, d_closest, d_block_pots, pot_stride);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_kpp_update_batched(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const double *d_lut768, const double *d_rows,
                                     const double *d_cands, const int *d_pick, int first, double *d_closest,
                                     double *d_tile_sums, int n_batch, void *stream) {
	CS_REQUIRE(ctx && (d_rows || (d_px && d_lut768)) && d_cands && d_closest && d_tile_sums, "null pointer");
	CS_REQUIRE(n > 0 && n_batch >= 1 && n_batch <= 65535, "bad n or n_batch");
	const long long ntiles = (n + kTile - 1) / kTile;
	kpp_update_batched_kernel<<<dim3((unsigned)ntiles, n_batch), kThreads, 0, CS_STREAM>>>(
	    reinterpret_cast<const uint32_t *>(d_px), n, d_lut768, d_rows, d_cands, d_pick, first, d_closest, d_tile_sums);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_kpp_locate_batched(cs_ctx *ctx, const double *d_closest, int64_t n, const int64_t *d_tile,
                                     const double *d_prefix_val, int n_query, const uint8_t *d_px, int64_t *d_index,
                                     uint8_t *d_index_px, int n_batch, void *stream) {
	CS_REQUIRE(ctx && d_closest && d_tile && d_prefix_val && d_index, "null pointer");
	CS_REQUIRE(!d_index_px || d_px, "d_index_px needs d_px");
	CS_REQUIRE(n > 0 && n_query >= 1 && n_query <= kMaxTrials && n_batch >= 1, "n must be > 0, 1 <= n_query <= 8, n_batch >= 1");
	kpp_locate_batched_kernel<<<n_batch, 32, 0, CS_STREAM>>>(d_closest, n, reinterpret_cast<const long long *>(d_tile), d_prefix_val,
	                                                         n_query, reinterpret_cast<const uint32_t *>(d_px),
	                                                         reinterpret_cast<long long *>(d_index),
	                                                         reinterpret_cast<uint32_t *>(d_index_px));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_kpp_pick_batched(cs_ctx *ctx, const double *d_block_pots, int pot_stride, int n_blocks, int n_cand,
                                   int n_batch, int *d_pick, double *d_pot, void *stream) {
	CS_REQUIRE(ctx && d_block_pots && d_pick && d_pot, "null pointer");
	CS_REQUIRE(n_blocks >= 1 && n_blocks <= pot_stride && n_cand >= 1 && n_cand <= kMaxTrials && n_batch >= 1, "bad sizes");
	kpp_pick_batched_kernel<<<n_batch, 32, 0, CS_STREAM>>>(d_block_pots, pot_stride, n_blocks, n_cand, d_pick, d_pot);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_kpp_draw_batched(cs_ctx *ctx, const double *d_closest, int64_t n, const double *d_tile_sums, const double *d_pot,
                                   const double *d_uniforms, int n_query, const uint8_t *d_px, const double *d_lut768,
                                   const double *d_rows, int64_t *d_cand_index, double *d_cands, int n_batch, void *stream) {
	CS_REQUIRE(ctx && d_closest && d_tile_sums && d_pot && d_uniforms && d_cand_index && d_cands, "null pointer");
	CS_REQUIRE(d_rows || (d_px && d_lut768), "samples: d_rows, or d_px with d_lut768");
	CS_REQUIRE(n > 0 && n_query >= 1 && n_query <= kMaxTrials && n_batch >= 1, "n must be > 0, 1 <= n_query <= 8, n_batch >= 1");
	kpp_draw_batched_kernel<<<n_batch, 32 * n_query, 0, CS_STREAM>>>(d_closest, n, d_tile_sums, (n + kTile - 1) / kTile, d_pot, d_uniforms, n_query,
	                                                       reinterpret_cast<const uint32_t *>(d_px), d_lut768, d_rows,
	                                                       reinterpret_cast<long long *>(d_cand_index), d_cands);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_kpp_record_batched(cs_ctx *ctx, const int64_t *d_cand_index, const double *d_cands, const int *d_pick, int slot,
                                     int K, int n_batch, int64_t *d_index_out, double *d_centers_out, void *stream) {
	CS_REQUIRE(ctx && d_cand_index && d_cands && d_pick && d_index_out && d_centers_out, "null pointer");
	CS_REQUIRE(K >= 1 && slot >= 0 && slot < K && n_batch >= 1, "bad slot, K or n_batch");
	kpp_record_batched_kernel<<<(n_batch + 63) / 64, 64, 0, CS_STREAM>>>(reinterpret_cast<const long long *>(d_cand_index), d_cands, d_pick,
	                                                                     slot, K, n_batch, reinterpret_cast<long long *>(d_index_out),
	                                                                     d_centers_out);
	CS_CUDA(cudaGetLastError());
	return 0;
}
