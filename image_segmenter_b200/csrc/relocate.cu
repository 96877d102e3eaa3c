// relocate.cu — empty-cluster relocation (rare path of the M-step) and the host-buffer
// convenience entry point.
//
// cs_lloyd_relocate_* replaces _relocate_empty_clusters_dense
// (sklearn/cluster/_k_means_common.pyx:167-211).  It is only needed when a cluster received
// no pixel, so it favours simplicity: per empty cluster two reduction passes over the pixels
// (largest squared distance to the assigned old centre, then the lowest pixel index holding
// it) and a one-thread fix-up of the sums.  Synchronous on the given stream.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;

struct Feat {
	const float *f0, *f1, *f2;
	const uint32_t *rgba;
	const float *lut3;  // optional 3 x 256 per-byte feature tables for packed pixels
	__device__ __forceinline__ void get(long long i, double &x, double &y, double &z) const {
		if (rgba) {
			const uint32_t w = rgba[i];
			if (lut3) {
				x = (double)lut3[w & 0xFFu]; y = (double)lut3[256 + ((w >> 8) & 0xFFu)]; z = (double)lut3[512 + ((w >> 16) & 0xFFu)];
			} else {
				x = (double)(w & 0xFFu); y = (double)((w >> 8) & 0xFFu); z = (double)((w >> 16) & 0xFFu);
			}
		} else {
			x = (double)f0[i]; y = (double)f1[i]; z = (double)f2[i];
		}
	}
};

// ((X - C_old[labels])**2).sum(axis=1) as NumPy evaluates it: three rounded squares, added in order
__device__ __forceinline__ double dist_to_own(const Feat &f, long long i, const uint8_t *labels,
                                              const double *c_old) {
	double x, y, z;
	f.get(i, x, y, z);
	const int l = labels[i];
	const double dx = x - c_old[3 * l], dy = y - c_old[3 * l + 1], dz = z - c_old[3 * l + 2];
	return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// scratch layout (u64): [0] = best distance bits, [1] = best index, [2] = previous distance
// bits, [3] = previous index (~0 = none)
// `base` = global index of this shard's first pixel (0 when unsharded): picks are ordered by (distance
// descending, GLOBAL index ascending), so a row-sharded run visits the same points as an unsharded one.
__global__ void __launch_bounds__(kThreads) far_dist_kernel(Feat f, long long n, const uint8_t *labels,
                                                            const double *c_old, int K,
                                                            unsigned long long *scr, unsigned long long base) {
	const unsigned long long pd = scr[2], pi = scr[3];
	unsigned long long best = 0ull;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		if (labels[i] >= K) continue;
		const unsigned long long d = (unsigned long long)__double_as_longlong(dist_to_own(f, i, labels, c_old));
		const bool after_prev = pi == ~0ull || d < pd || (d == pd && base + (unsigned long long)i > pi);
		if (after_prev && d > best) best = d;
	}
	for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
	if ((threadIdx.x & 31) == 0 && best) atomicMax(scr + 0, best);
}
__global__ void __launch_bounds__(kThreads) far_index_kernel(Feat f, long long n, const uint8_t *labels,
                                                             const double *c_old, int K,
                                                             unsigned long long *scr, unsigned long long base) {
	const unsigned long long pd = scr[2], pi = scr[3], target = scr[0];
	unsigned long long best = ~0ull;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		if (labels[i] >= K) continue;
		const unsigned long long d = (unsigned long long)__double_as_longlong(dist_to_own(f, i, labels, c_old));
		const bool after_prev = pi == ~0ull || d < pd || (d == pd && base + (unsigned long long)i > pi);
		if (after_prev && d == target) best = min(best, (unsigned long long)i);
	}
	for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
	if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(scr + 1, best);
}
__global__ void relocate_apply_kernel(Feat f, const uint8_t *labels, int new_id, double *sums,
                                      double *counts, unsigned long long *scr) {
	const unsigned long long far = scr[1];
	if (far != ~0ull) {
		double x, y, z;
		f.get((long long)far, x, y, z);
		const int old_id = labels[far];
		sums[3 * old_id] -= x; sums[3 * old_id + 1] -= y; sums[3 * old_id + 2] -= z;
		sums[3 * new_id] = x; sums[3 * new_id + 1] = y; sums[3 * new_id + 2] = z;
		counts[new_id] = 1.0;
		counts[old_id] -= 1.0;
	}
	scr[2] = scr[0]; scr[3] = scr[1];  // this pick becomes "previous"
	scr[0] = 0ull; scr[1] = ~0ull;
}

// the local pick as a record for the host: {distance bits, global index (~0 = none), x, y, z, label}
__global__ void far_report_kernel(Feat f, const uint8_t *labels, unsigned long long *scr, unsigned long long base,
                                  unsigned long long *out6) {
	const unsigned long long far = scr[1];
	out6[0] = scr[0];
	out6[1] = far == ~0ull ? ~0ull : base + far;
	double x = 0.0, y = 0.0, z = 0.0;
	unsigned long long lab = 0ull;
	if (far != ~0ull) { f.get((long long)far, x, y, z); lab = labels[far]; }
	out6[2] = (unsigned long long)__double_as_longlong(x);
	out6[3] = (unsigned long long)__double_as_longlong(y);
	out6[4] = (unsigned long long)__double_as_longlong(z);
	out6[5] = lab;
}

int relocate_impl(cs_ctx *ctx, Feat f, int64_t n, const uint8_t *d_labels, const double *d_centers_old,
                  int K, double *d_sums, double *d_counts, cudaStream_t st) {
	double h_counts[CS_MAX_K];
	CS_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
	CS_CUDA(cudaStreamSynchronize(st));
	int empty[CS_MAX_K], n_empty = 0;
	for (int k = 0; k < K; ++k)
		if (h_counts[k] == 0.0) empty[n_empty++] = k;
	if (n_empty == 0 || n == 0) return 0;
	unsigned long long init[4] = {0ull, ~0ull, 0ull, ~0ull};
	unsigned long long *scr = ctx->d_scratch64;
	CS_CUDA(cudaMemcpyAsync(scr, init, sizeof(init), cudaMemcpyHostToDevice, st));
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	for (int e = 0; e < n_empty; ++e) {
		far_dist_kernel<<<grid, kThreads, 0, st>>>(f, n, d_labels, d_centers_old, K, scr, 0ull);
		if (e == 0) {
			// np.max(distances) == 0  ->  relocation is pointless, sklearn returns early
			unsigned long long top;
			CS_CUDA(cudaMemcpyAsync(&top, scr, sizeof(top), cudaMemcpyDeviceToHost, st));
			CS_CUDA(cudaStreamSynchronize(st));
			if (top == 0ull) return 0;
		}
		far_index_kernel<<<grid, kThreads, 0, st>>>(f, n, d_labels, d_centers_old, K, scr, 0ull);
		relocate_apply_kernel<<<1, 1, 0, st>>>(f, d_labels, empty[e], d_sums, d_counts, scr);
	}
	CS_CUDA(cudaGetLastError());
	CS_CUDA(cudaStreamSynchronize(st));
	return 0;
}

} // namespace
} // namespace cs

using namespace cs;

extern "C" int cs_lloyd_relocate_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                                     int64_t n, const uint8_t *d_labels, const double *d_centers_old, int K,
                                     double *d_sums, double *d_counts, void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_labels && d_centers_old && d_sums && d_counts, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0, "bad K or n");
	Feat f{d_f0, d_f1, d_f2, nullptr, nullptr};
	return relocate_impl(ctx, f, n, d_labels, d_centers_old, K, d_sums, d_counts, (cudaStream_t)stream);
}

extern "C" int cs_lloyd_relocate_px8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, const float *d_lut3,
                                     const uint8_t *d_labels, const double *d_centers_old, int K, double *d_sums,
                                     double *d_counts, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_labels && d_centers_old && d_sums && d_counts, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0, "bad K or n");
	Feat f{nullptr, nullptr, nullptr, reinterpret_cast<const uint32_t *>(d_rgba), d_lut3};
	return relocate_impl(ctx, f, n, d_labels, d_centers_old, K, d_sums, d_counts, (cudaStream_t)stream);
}

// One step of a row-SHARDED relocation (image_segmenter_b200/sharded.py): this rank's farthest labelled
// pixel that comes after the previous global pick (h_prev2 = {distance bits, global index}, index ~0 = no
// previous pick) in the order (distance to its own old centre descending, global index ascending).
// h_out6 = {distance bits, global index (~0 = this shard has none), x, y, z as double bits, label}.
// The ranks all-gather their records, take the first in that order and apply it to the (already
// all-reduced) sums / counts exactly as relocate_apply_kernel does.  Synchronous.
extern "C" int cs_lloyd_farthest_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                                     int64_t n, const uint8_t *d_labels, const double *d_centers_old, int K,
                                     uint64_t index_base, const uint64_t *h_prev2, uint64_t *h_out6, void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_labels && d_centers_old && h_prev2 && h_out6, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0, "bad K or n");
	cudaStream_t st = (cudaStream_t)stream;
	Feat f{d_f0, d_f1, d_f2, nullptr, nullptr};
	unsigned long long init[4] = {0ull, ~0ull, h_prev2[0], h_prev2[1]};
	unsigned long long *scr = ctx->d_scratch64;
	CS_CUDA(cudaMemcpyAsync(scr, init, sizeof(init), cudaMemcpyHostToDevice, st));
	if (n > 0) {
		const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
		far_dist_kernel<<<grid, kThreads, 0, st>>>(f, n, d_labels, d_centers_old, K, scr, index_base);
		far_index_kernel<<<grid, kThreads, 0, st>>>(f, n, d_labels, d_centers_old, K, scr, index_base);
	}
	far_report_kernel<<<1, 1, 0, st>>>(f, d_labels, scr, index_base, scr + 8);
	CS_CUDA(cudaGetLastError());
	CS_CUDA(cudaMemcpyAsync(h_out6, scr + 8, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CS_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// ---- host-buffer convenience: upload, convert, iterate, download ----------------------------
extern "C" int cs_host_lab_kmeans(cs_ctx *ctx, const uint8_t *h_rgba, int64_t n, const double *h_lut256,
                                  double *h_centers, int K, int n_iter, double tol, int flags,
                                  uint8_t *h_labels, int *n_iter_done, double *h_inertia) {
	CS_REQUIRE(ctx && h_rgba && h_lut256 && h_centers, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n > 0 && n_iter >= 0, "bad K, n or n_iter");
	CS_CUDA(cudaSetDevice(ctx->device));
	// the call runs on its own two non-blocking streams but shares the context's scratch (partials, block
	// counters) with everything queued earlier through this context: wait for that work first
	CS_CUDA(cudaDeviceSynchronize());
	const size_t n4 = ((size_t)n + 3) & ~(size_t)3;
	// layout: rgba | L | a | b | labels | doubles (lut 256, centres 2 x 3K, sums 3K, counts K, stats 4, inertia 1)
	const size_t off_L = n4 * 4, off_a = off_L + n4 * 4, off_b = off_a + n4 * 4, off_lab = off_b + n4 * 4;
	const size_t off_d = (off_lab + n4 + 255) & ~(size_t)255;
	const size_t n_dbl = 256 + 6 * (size_t)K + 3 * (size_t)K + K + 4 + 1 + 4;
	const size_t need = off_d + n_dbl * sizeof(double);
	if (ctx->host_buf_bytes < need) {
		if (ctx->d_host_buf) cudaFree(ctx->d_host_buf);
		ctx->d_host_buf = nullptr; ctx->host_buf_bytes = 0;
		CS_CUDA(cudaMalloc(&ctx->d_host_buf, need));
		ctx->host_buf_bytes = need;
	}
	unsigned char *base = static_cast<unsigned char *>(ctx->d_host_buf);
	uint8_t *d_rgba = base;
	float *d_L = reinterpret_cast<float *>(base + off_L), *d_a = reinterpret_cast<float *>(base + off_a),
	      *d_b = reinterpret_cast<float *>(base + off_b);
	uint8_t *d_lab = base + off_lab;
	double *d_lut = reinterpret_cast<double *>(base + off_d);
	double *d_c[2] = {d_lut + 256, d_lut + 256 + 3 * K};
	double *d_sums = d_lut + 256 + 6 * K, *d_counts = d_sums + 3 * K, *d_stats = d_counts + K, *d_inert = d_stats + 4;
	if (!ctx->host_copy) {
		CS_CUDA(cudaStreamCreateWithFlags(&ctx->host_copy, cudaStreamNonBlocking));
		CS_CUDA(cudaStreamCreateWithFlags(&ctx->host_comp, cudaStreamNonBlocking));
		for (cudaEvent_t &e : ctx->host_ev) CS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	cudaStream_t st = ctx->host_comp, cp = ctx->host_copy;
	CS_CUDA(cudaMemcpyAsync(d_lut, h_lut256, 256 * sizeof(double), cudaMemcpyHostToDevice, st));
	CS_CUDA(cudaMemcpyAsync(d_c[0], h_centers, sizeof(double) * 3 * K, cudaMemcpyHostToDevice, st));
	// the image goes up in up to 8 chunks on the copy stream; the LAB conversion of chunk i runs on the
	// compute stream while chunk i+1 is still on the bus (chunk sizes are multiples of 4 px: 16-byte planes)
	int rc = 0;
	{
		constexpr int kChunks = 8;
		const int64_t per = ((n + kChunks - 1) / kChunks + 3) & ~(int64_t)3;
		int ci = 0;
		for (int64_t o = 0; o < n; o += per, ++ci) {
			const int64_t m = n - o < per ? n - o : per;
			CS_CUDA(cudaMemcpyAsync(d_rgba + o * 4, h_rgba + o * 4, (size_t)m * 4, cudaMemcpyHostToDevice, cp));
			CS_CUDA(cudaEventRecord(ctx->host_ev[ci], cp));
			CS_CUDA(cudaStreamWaitEvent(st, ctx->host_ev[ci], 0));
			rc = cs_rgba8_to_lab(ctx, d_rgba + o * 4, m, d_lut, d_L + o, d_a + o, d_b + o, st);
			if (rc) return rc;
		}
	}
	// iterations are queued in batches with the convergence / empty-cluster test on the device
	// (cs_lloyd_run_f32): one host round trip per batch instead of one per iteration
	double *d_ctl = d_inert + 1;
	double ctl[4] = {0.0, 0.0, tol, 0.0};
	CS_CUDA(cudaMemcpyAsync(d_ctl, ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
	int cur = 0, it = 0;
	while (it < n_iter) {
		const int m = n_iter - it < 10 ? n_iter - it : 10;
		rc = cs_lloyd_run_f32(ctx, d_L, d_a, d_b, n, d_c[cur], d_c[cur ^ 1], K, d_sums, d_counts, d_stats, CS_LAB_NORM2_MAX,
		                      flags, m, d_ctl, st);
		if (rc) return rc;
		CS_CUDA(cudaMemcpyAsync(ctl, d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
		CS_CUDA(cudaStreamSynchronize(st));
		const int done = (int)ctl[1] - it;
		cur ^= done & 1;
		it += done;
		if (ctl[0] == 1.0) break;  // covers sklearn's strict (labels unchanged => shift 0) and tol stops
		if (ctl[0] == 2.0) {
			// an empty cluster: redo the step with labels, relocate, finish the M-step
			rc = cs_lloyd_step_f32(ctx, d_L, d_a, d_b, n, d_c[cur], K, d_lab, d_sums, d_counts, nullptr,
			                       CS_LAB_NORM2_MAX, flags, st);
			if (rc) return rc;
			rc = cs_lloyd_relocate_f32(ctx, d_L, d_a, d_b, n, d_lab, d_c[cur], K, d_sums, d_counts, st);
			if (rc) return rc;
			rc = cs_lloyd_finalize(ctx, d_sums, d_counts, d_c[cur], K, d_c[cur ^ 1], d_stats, st);
			if (rc) return rc;
			double stats[4];
			CS_CUDA(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
			CS_CUDA(cudaStreamSynchronize(st));
			cur ^= 1;
			++it;
			ctl[0] = 0.0; ctl[1] = (double)it; ctl[2] = tol;
			CS_CUDA(cudaMemcpyAsync(d_ctl, ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
			if (stats[0] <= tol) break;
		}
	}
	// final E-step on the final centres: labels (+ inertia)
	rc = cs_lloyd_step_f32(ctx, d_L, d_a, d_b, n, d_c[cur], K, d_lab, d_sums, d_counts, d_inert, CS_LAB_NORM2_MAX,
	                       flags, st);
	if (rc) return rc;
	CS_CUDA(cudaMemcpyAsync(h_centers, d_c[cur], sizeof(double) * 3 * K, cudaMemcpyDeviceToHost, st));
	if (h_labels) CS_CUDA(cudaMemcpyAsync(h_labels, d_lab, (size_t)n, cudaMemcpyDeviceToHost, st));
	if (h_inertia) CS_CUDA(cudaMemcpyAsync(h_inertia, d_inert, sizeof(double), cudaMemcpyDeviceToHost, st));
	CS_CUDA(cudaStreamSynchronize(st));
	if (n_iter_done) *n_iter_done = it;
	return 0;
}
