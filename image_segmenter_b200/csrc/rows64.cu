// rows64.cu — the fp64-row kernels behind simplify_colors_adaptive_distance, sm_100a.
//
// That entry point (app/processing/color_simplify.py:710-882) works on the fp64 CIELAB rows of every opaque
// pixel (skimage rgb2lab -> StandardScaler).  Two of its steps are per-pixel work over those rows:
//   :809-814  KMeans(n_clusters=num_colors, random_state=42, n_init=10).fit_predict(lab_normalized)   (fallback when
//             DBSCAN found fewer clusters than asked for)  -> cs_lloyd_step_rows64 + cs_sum_by_label_rows64
//             (+ cs_lloyd_finalize and the k-means++ passes of kmeanspp.cu on rows)
//   :861-867  pairwise_distances_argmin_min(lab_flat[dark], lab_filtered)                             -> cs_nn_argmin_rows64
// fp64 throughout (the reference's dtype), direct squared distances, strict `<` (first minimum), sums in a fixed
// order (run-to-run deterministic).  These are small-image kernels (DBSCAN itself limits the function to ~10^5
// pixels): clarity over speed, but every loop over pixels is on the device.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;
constexpr int kRefTile = 1024;  // reference rows staged in shared memory per step (24 KB)

__device__ __forceinline__ double d2(const double *a, double x, double y, double z) {
	const double dx = x - a[0], dy = y - a[1], dz = z - a[2];
	return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// index of the nearest reference row for every query row (first minimum)
__global__ void __launch_bounds__(kThreads) nn_argmin_kernel(const double *__restrict__ q, long long nq,
                                                             const double *__restrict__ ref, long long nr,
                                                             long long *__restrict__ out) {
	__shared__ double tile[kRefTile * 3];
	const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
	const bool live = i < nq;
	const double x = live ? q[3 * i] : 0.0, y = live ? q[3 * i + 1] : 0.0, z = live ? q[3 * i + 2] : 0.0;
	double best = 1e300;
	long long bi = 0;
	for (long long r0 = 0; r0 < nr; r0 += kRefTile) {
		const int m = (int)(nr - r0 < kRefTile ? nr - r0 : kRefTile);
		__syncthreads();
		for (int t = threadIdx.x; t < m * 3; t += kThreads) tile[t] = ref[3 * r0 + t];
		__syncthreads();
		if (live)
			for (int r = 0; r < m; ++r) {
				const double d = d2(tile + 3 * r, x, y, z);
				if (d < best) { best = d; bi = r0 + r; }
			}
	}
	if (live) out[i] = bi;
}

// E-step on fp64 rows: labels (int32) + per-block partial inertia (summed in block order by the caller's second launch)
__global__ void __launch_bounds__(kThreads) step_rows64_kernel(const double *__restrict__ rows, long long n,
                                                               const double *__restrict__ centers, int K,
                                                               int *__restrict__ labels, double *__restrict__ block_inertia) {
	__shared__ double c[CS_MAX_K * 3];
	__shared__ double red[kThreads];
	for (int t = threadIdx.x; t < K * 3; t += kThreads) c[t] = centers[t];
	__syncthreads();
	double acc = 0.0;
	const long long per = (n + gridDim.x - 1) / gridDim.x;  // contiguous chunk per block: fixed assignment
	const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
	for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
		const double x = rows[3 * i], y = rows[3 * i + 1], z = rows[3 * i + 2];
		double best = 1e300;
		int bk = 0;
		for (int k = 0; k < K; ++k) {
			const double d = d2(c + 3 * k, x, y, z);
			if (d < best) { best = d; bk = k; }
		}
		labels[i] = bk;
		acc += best;
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int o = kThreads / 2; o > 0; o >>= 1) {
		if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) block_inertia[blockIdx.x] = red[0];
}

__global__ void sum_blocks_kernel(const double *__restrict__ v, int nb, double *__restrict__ out) {
	double s = 0.0;
	for (int b = 0; b < nb; ++b) s += v[b];
	*out = s;
}

// block k: sum of the rows with label k, fixed order (thread-strided partials, then a fixed tree)
__global__ void __launch_bounds__(kThreads) sum_by_label_rows64_kernel(const double *__restrict__ rows, long long n,
                                                                       const int *__restrict__ labels,
                                                                       double *__restrict__ sums, double *__restrict__ counts) {
	__shared__ double red[4][kThreads];
	const int k = blockIdx.x;
	double s0 = 0.0, s1 = 0.0, s2 = 0.0, cn = 0.0;
	for (long long i = threadIdx.x; i < n; i += kThreads)
		if (labels[i] == k) { s0 += rows[3 * i]; s1 += rows[3 * i + 1]; s2 += rows[3 * i + 2]; cn += 1.0; }
	red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2; red[3][threadIdx.x] = cn;
	__syncthreads();
	for (int o = kThreads / 2; o > 0; o >>= 1) {
		if (threadIdx.x < o)
			for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) { sums[3 * k] = red[0][0]; sums[3 * k + 1] = red[1][0]; sums[3 * k + 2] = red[2][0]; counts[k] = red[3][0]; }
}

} // namespace
} // namespace cs

using namespace cs;
#define CS_STREAM ((cudaStream_t)stream)

extern "C" int cs_nn_argmin_rows64(cs_ctx *ctx, const double *d_query, int64_t n_query, const double *d_ref, int64_t n_ref,
                                   int64_t *d_index, void *stream) {
	CS_REQUIRE(ctx && d_query && d_ref && d_index, "null pointer");
	CS_REQUIRE(n_query >= 0 && n_ref >= 1, "n_query must be >= 0 and n_ref >= 1");
	if (n_query == 0) return 0;
	nn_argmin_kernel<<<(unsigned)((n_query + kThreads - 1) / kThreads), kThreads, 0, CS_STREAM>>>(
	    d_query, n_query, d_ref, n_ref, reinterpret_cast<long long *>(d_index));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_lloyd_step_rows64(cs_ctx *ctx, const double *d_rows, int64_t n, const double *d_centers, int K,
                                    int32_t *d_labels, double *d_inertia, void *stream) {
	CS_REQUIRE(ctx && d_rows && d_centers && d_labels && d_inertia, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 1, "bad K or n");
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 4);
	double *blk = ctx->d_partials;  // scratch: one partial per block (grid <= 4 * SMs << kMaxPartialBlocks * kMaxPartialVals)
	step_rows64_kernel<<<grid, kThreads, 0, CS_STREAM>>>(d_rows, n, d_centers, K, d_labels, blk);
	sum_blocks_kernel<<<1, 1, 0, CS_STREAM>>>(blk, grid, d_inertia);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_sum_by_label_rows64(cs_ctx *ctx, const double *d_rows, int64_t n, const int32_t *d_labels, int K,
                                      double *d_sums, double *d_counts, void *stream) {
	CS_REQUIRE(ctx && d_rows && d_labels && d_sums && d_counts, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0, "bad K or n");
	sum_by_label_rows64_kernel<<<K, kThreads, 0, CS_STREAM>>>(d_rows, n, d_labels, d_sums, d_counts);
	CS_CUDA(cudaGetLastError());
	return 0;
}
