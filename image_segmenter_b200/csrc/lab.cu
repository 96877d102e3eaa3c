// lab.cu — K1 (sRGB u8 -> CIELAB), K9 (RGB -> OpenCV 8-bit HSV) and K4 (nearest centre +
// palette remap, fused with the colour-space conversion), sm_100a.
//
// K1 replaces skimage.color.rgb2lab (app/processing/color_simplify.py:470, 540, 658, 688, 757,
// 1090-1091); K9 replaces cv2.cvtColor(COLOR_RGB2HSV) on uint8 (:947, 1097-1098); K4 replaces
// sklearn pairwise_distances_argmin_min + the gather/alpha/dstack epilogue (:543-557, 691-705,
// 1106-1121).  All three are streaming kernels: 16-byte loads of 4 RGBA8 pixels per thread,
// conversion in registers, 16-byte stores.
#include "cs_common.cuh"

namespace cs {
namespace {

constexpr int kThreads = 256;

// skimage.color.colorconv.xyz_from_rgb and the D65 / 2 degree white point
__constant__ double kM[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160,
                             0.072169, 0.019334, 0.119193, 0.950227};
__constant__ double kWhite[3] = {0.95047, 1.0, 1.08883};
__constant__ double kWhiteInv[3] = {1.0 / 0.95047, 1.0, 1.0 / 1.08883};

// cube root to ~1e-14 relative: fp32 cbrtf seed (1 ulp of fp32) + one Newton step whose residual
// t - y^3 is formed in fp64 with a single rounding; the step's 1/(3 y^2) only needs fp32 accuracy.
__device__ __forceinline__ double cbrt_fast(double t) {
	const float yf = cbrtf((float)t);
	const double y = (double)yf;
	const double r = fma(-(y * y), y, t);  // y*y is exact in fp64 (24-bit y)
	return fma(r, (double)__frcp_rn(3.f * yf * yf), y);
}

// one pixel, fp64, same operation order as the NumPy expression of skimage's rgb2xyz/xyz2lab
// (explicit _rn intrinsics keep nvcc from contracting mul+add pairs NumPy rounds separately).
// EXACT = true: IEEE division by the white point and libm cbrt (the fp64 rows handed to DBSCAN);
// EXACT = false: reciprocal multiply and cbrt_fast — differs from the exact path by ~1e-14 relative,
// far below the single rounding to fp32 that follows (K1) and below any distance gap that matters (K4).
template <bool EXACT>
__device__ __forceinline__ void rgb_to_lab_f64(const double *lut, uint32_t r, uint32_t g,
                                               uint32_t b, double &L, double &A, double &B) {
	const double lr = lut[r], lg = lut[g], lb = lut[b];
	double f[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		double v = __dadd_rn(__dadd_rn(__dmul_rn(lr, kM[3 * i]), __dmul_rn(lg, kM[3 * i + 1])),
		                     __dmul_rn(lb, kM[3 * i + 2]));
		v = EXACT ? v / kWhite[i] : v * kWhiteInv[i];
		f[i] = v > 0.008856 ? (EXACT ? cbrt(v) : cbrt_fast(v)) : __dadd_rn(__dmul_rn(7.787, v), 16.0 / 116.0);
	}
	L = __dsub_rn(__dmul_rn(116.0, f[1]), 16.0);
	A = __dmul_rn(500.0, __dsub_rn(f[0], f[1]));
	B = __dmul_rn(200.0, __dsub_rn(f[1], f[2]));
}

// OpenCV RGB2HSV_b (imgproc/src/color_hsv.simd.hpp): 12-bit fixed-point reciprocal tables,
// sdiv[i] = cvRound((255<<12)/i), hdiv[i] = cvRound((180<<12)/(6 i)); tables in shared memory.
__device__ __forceinline__ void rgb_to_hsv_u8(const int *sdiv, const int *hdiv, int r, int g, int b,
                                              int &h, int &s, int &v) {
	v = max(max(r, g), b);
	const int vmin = min(min(r, g), b);
	const int diff = v - vmin;
	s = (diff * sdiv[v] + (1 << 11)) >> 12;
	int hn = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
	h = (hn * hdiv[diff] + (1 << 11)) >> 12;  // arithmetic shift of a possibly negative value
	if (h < 0) h += 180;
	h = min(max(h, 0), 255); s = min(max(s, 0), 255);
}

__device__ __forceinline__ void fill_hsv_tables(int *sdiv, int *hdiv) {
	for (int i = threadIdx.x; i < 256; i += blockDim.x) {
		// rint == cvRound (round half to even)
		sdiv[i] = i ? (int)rint((double)(255 << 12) / (double)i) : 0;
		hdiv[i] = i ? (int)rint((double)(180 << 12) / (6.0 * (double)i)) : 0;
	}
}

__global__ void __launch_bounds__(kThreads) rgba8_to_lab_kernel(
    const uint32_t *__restrict__ rgba, long long n, const double *__restrict__ lut_g,
    float *__restrict__ oL, float *__restrict__ oA, float *__restrict__ oB) {
	__shared__ double lut[256];
	for (int i = threadIdx.x; i < 256; i += kThreads) lut[i] = lut_g[i];
	__syncthreads();
	const long long n4 = n >> 2;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
		const uint4 px = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i);
		const uint32_t w[4] = {px.x, px.y, px.z, px.w};
		float l[4], a[4], b[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			double L, A, B;
			rgb_to_lab_f64<false>(lut, w[q] & 0xFFu, (w[q] >> 8) & 0xFFu, (w[q] >> 16) & 0xFFu, L, A, B);
			l[q] = (float)L; a[q] = (float)A; b[q] = (float)B;
		}
		reinterpret_cast<float4 *>(oL)[i] = make_float4(l[0], l[1], l[2], l[3]);
		reinterpret_cast<float4 *>(oA)[i] = make_float4(a[0], a[1], a[2], a[3]);
		reinterpret_cast<float4 *>(oB)[i] = make_float4(b[0], b[1], b[2], b[3]);
	}
	// ragged tail (n % 4 pixels)
	if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
		const long long i = (n4 << 2) + threadIdx.x;
		const uint32_t w = rgba[i];
		double L, A, B;
		rgb_to_lab_f64<false>(lut, w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, L, A, B);
		oL[i] = (float)L; oA[i] = (float)A; oB[i] = (float)B;
	}
}

// fp64 rows (n x 3, interleaved): the array skimage hands to DBSCAN / StandardScaler in
// simplify_colors_adaptive_distance (color_simplify.py:757)
__global__ void __launch_bounds__(kThreads) rgba8_to_lab_f64_kernel(const uint32_t *__restrict__ rgba, long long n,
                                                                    const double *__restrict__ lut_g,
                                                                    double *__restrict__ out) {
	__shared__ double lut[256];
	for (int i = threadIdx.x; i < 256; i += kThreads) lut[i] = lut_g[i];
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t w = rgba[i];
		double L, A, B;
		rgb_to_lab_f64<true>(lut, w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, L, A, B);
		out[3 * i] = L; out[3 * i + 1] = A; out[3 * i + 2] = B;
	}
}

__device__ __forceinline__ uint32_t hsv_word(const int *sdiv, const int *hdiv, uint32_t w) {
	int h, s, v;
	rgb_to_hsv_u8(sdiv, hdiv, w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF, h, s, v);
	return (uint32_t)h | ((uint32_t)s << 8) | ((uint32_t)v << 16) | (w & 0xFF000000u);
}

// 4 pixels per thread: one 16-byte streaming load and store; a scalar tail for n % 4 / unaligned buffers
__global__ void __launch_bounds__(kThreads) rgba8_to_hsv8_kernel(const uint32_t *__restrict__ rgba,
                                                                 long long n,
                                                                 uint32_t *__restrict__ out, int vec_ok) {
	__shared__ int sdiv[256], hdiv[256];
	fill_hsv_tables(sdiv, hdiv);
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	const long long n4 = vec_ok ? (n >> 2) : 0;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += 2 * stride) {
		const bool two = i + stride < n4;  // two 16-byte loads in flight per thread
		uint4 px = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i);
		uint4 py = two ? ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba) + i + stride) : make_uint4(0u, 0u, 0u, 0u);
		px.x = hsv_word(sdiv, hdiv, px.x); px.y = hsv_word(sdiv, hdiv, px.y);
		px.z = hsv_word(sdiv, hdiv, px.z); px.w = hsv_word(sdiv, hdiv, px.w);
		stg_stream_u4(reinterpret_cast<uint4 *>(out) + i, px);
		if (two) {
			py.x = hsv_word(sdiv, hdiv, py.x); py.y = hsv_word(sdiv, hdiv, py.y);
			py.z = hsv_word(sdiv, hdiv, py.z); py.w = hsv_word(sdiv, hdiv, py.w);
			stg_stream_u4(reinterpret_cast<uint4 *>(out) + i + stride, py);
		}
	}
	for (long long i = (n4 << 2) + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
		out[i] = hsv_word(sdiv, hdiv, rgba[i]);
}

__global__ void __launch_bounds__(kThreads) norm2_max_kernel(const float *__restrict__ f0,
                                                             const float *__restrict__ f1,
                                                             const float *__restrict__ f2,
                                                             long long n, unsigned long long *out) {
	float m = 0.f;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const float a = f0[i], b = f1[i], c = f2[i];
		m = fmaxf(m, fmaf(a, a, fmaf(b, b, c * c)));
	}
	for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
	// non-negative doubles order like their bit patterns
	if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong((double)m * 1.000001));
}

// K4.  SPACE: 0 RGB, 1 LAB (fp64, not rounded to fp32 — the reference's LAB is fp64), 2 HSV u8.
// The pixel's features are formed exactly (fp64); every centre is first screened in fp32 with
// v_k = |c_k|^2 - 2 x.c_k (3 FFMA), tracking the smallest and second smallest value; only when the two
// are closer than the fp32 error bound of v is the pixel re-evaluated with the direct fp64 formula and
// strict `<` — so the label always equals the fp64 first-minimum (lowest index on exact ties,
// sklearn/utils/_heap.pyx:46-47) at a fraction of the fp64 work.
struct RemapTab {
	float4 f32[CS_MAX_K];      // {-2cx, -2cy, -2cz, |c|^2}
	double c[CS_MAX_K * 3];
	uint32_t pal[CS_MAX_K];
	float cmax;                // max_k |c_k|
};

__device__ __noinline__ int remap_exact_label(double x, double y, double z, const double *c, int K) {
	double best = 1e300;
	int bi = 0;
	for (int k = 0; k < K; ++k) {  // strict < : lowest index wins exact ties
		const double dx = x - c[3 * k], dy = y - c[3 * k + 1], dz = z - c[3 * k + 2];
		const double d = dx * dx + dy * dy + dz * dz;
		if (d < best) { best = d; bi = k; }
	}
	return bi;
}

template <int SPACE>
__global__ void __launch_bounds__(kThreads) assign_remap_kernel(
    const uint32_t *__restrict__ rgba, long long n, const double *__restrict__ lut_g,
    const double *__restrict__ centers, const uint8_t *__restrict__ palette, int K,
    int preserve_alpha, uint32_t *__restrict__ out, uint8_t *__restrict__ labels) {
	__shared__ double lut[SPACE == 1 ? 256 : 1];
	__shared__ int sdiv[SPACE == 2 ? 256 : 1], hdiv[SPACE == 2 ? 256 : 1];
	__shared__ RemapTab T;
	if (SPACE == 1)
		for (int i = threadIdx.x; i < 256; i += kThreads) lut[i] = lut_g[i];
	if (SPACE == 2) fill_hsv_tables(sdiv, hdiv);
	for (int i = threadIdx.x; i < K * 3; i += kThreads) T.c[i] = centers[i];
	for (int i = threadIdx.x; i < K; i += kThreads) {
		T.pal[i] = (uint32_t)palette[3 * i] | ((uint32_t)palette[3 * i + 1] << 8) | ((uint32_t)palette[3 * i + 2] << 16);
		const double cx = centers[3 * i], cy = centers[3 * i + 1], cz = centers[3 * i + 2];
		T.f32[i] = make_float4((float)(-2.0 * cx), (float)(-2.0 * cy), (float)(-2.0 * cz), (float)(cx * cx + cy * cy + cz * cz));
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		float m = 0.f;
		for (int k = 0; k < K; ++k) m = fmaxf(m, T.f32[k].w);
		T.cmax = sqrtf(m);
	}
	__syncthreads();
	const float cmax = T.cmax;
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t w = rgba[i];
		const uint32_t a = w >> 24;
		const uint32_t a_out = preserve_alpha ? a : (a > 128u ? 255u : 0u);
		uint32_t rgb_out = 0u;
		int best_k = 255;
		if (a > 0u) {
			const int r = w & 0xFF, g = (w >> 8) & 0xFF, b = (w >> 16) & 0xFF;
			double x, y, z;
			if (SPACE == 0) { x = r; y = g; z = b; }
			if (SPACE == 1) rgb_to_lab_f64<false>(lut, r, g, b, x, y, z);
			if (SPACE == 2) {
				int h, s, v;
				rgb_to_hsv_u8(sdiv, hdiv, r, g, b, h, s, v);
				x = h; y = s; z = v;
			}
			const float xf = (float)x, yf = (float)y, zf = (float)z;
			float best = 3.0e38f, second = 3.0e38f;
			best_k = 0;
#pragma unroll 4
			for (int k = 0; k < K; ++k) {
				const float4 t = T.f32[k];
				const float v = fmaf(xf, t.x, fmaf(yf, t.y, fmaf(zf, t.z, t.w)));
				const bool lt = v < best;
				second = lt ? best : fminf(second, v);
				best_k = lt ? k : best_k;
				best = fminf(best, v);
			}
			// |v_fp32 - v| <= 8 * 2^-24 * (|x| + max|c|)^2 covers the rounding of x, of the table and of the 3 FMAs
			const float xn = sqrtf(fmaf(xf, xf, fmaf(yf, yf, zf * zf))) + cmax;
			const float tau = 2.f * 8.f * 5.9604645e-8f * xn * xn;
			if (second - best <= tau) best_k = remap_exact_label(x, y, z, T.c, K);
			rgb_out = T.pal[best_k];
		}
		out[i] = rgb_out | (a_out << 24);
		if (labels) labels[i] = (uint8_t)best_k;
	}
}

__global__ void __launch_bounds__(kThreads) remap_labels_kernel(
    const uint32_t *__restrict__ rgba, const uint8_t *__restrict__ labels, long long n,
    const uint32_t *__restrict__ selpx, int mask_mode, int min_bright,
    const uint8_t *__restrict__ palette, int K, int preserve_alpha, uint32_t *__restrict__ out) {
	__shared__ uint32_t pal[CS_MAX_K];
	for (int i = threadIdx.x; i < CS_MAX_K; i += kThreads)
		pal[i] = i < K ? ((uint32_t)palette[3 * i] | ((uint32_t)palette[3 * i + 1] << 8) | ((uint32_t)palette[3 * i + 2] << 16)) : 0u;
	__syncthreads();
	const long long stride = (long long)gridDim.x * kThreads;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
		const uint32_t a = rgba[i] >> 24;
		const uint32_t a_out = preserve_alpha ? a : (a > 128u ? 255u : 0u);
		const uint32_t l = labels[i];
		const bool ok = a > 0u && l < (uint32_t)K && label_valid(selpx, i, mask_mode, min_bright, l, K);
		out[i] = (ok ? pal[l] : 0u) | (a_out << 24);
	}
}

} // namespace
} // namespace cs

using namespace cs;

extern "C" int cs_rgba8_to_lab(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, const double *d_lut256,
                               float *d_L, float *d_a, float *d_b, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_lut256 && d_L && d_a && d_b, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(((uintptr_t)d_rgba & 15u) == 0 && ((uintptr_t)d_L & 15u) == 0 && ((uintptr_t)d_a & 15u) == 0 &&
	               ((uintptr_t)d_b & 15u) == 0, "buffers must be 16-byte aligned");
	if (n == 0) return 0;
	const int grid = grid_for(ctx, (n / 4 + kThreads - 1) / kThreads, 8);
	rgba8_to_lab_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
	    reinterpret_cast<const uint32_t *>(d_rgba), n, d_lut256, d_L, d_a, d_b);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_rgba8_to_lab_f64(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, const double *d_lut256,
                                   double *d_lab, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_lut256 && d_lab, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(((uintptr_t)d_rgba & 3u) == 0 && ((uintptr_t)d_lab & 7u) == 0, "buffers must be naturally aligned");
	if (n == 0) return 0;
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	rgba8_to_lab_f64_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(d_rgba), n,
	                                                                    d_lut256, d_lab);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_rgba8_to_hsv8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, uint8_t *d_hsva,
                                void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_hsva, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(((uintptr_t)d_rgba & 3u) == 0 && ((uintptr_t)d_hsva & 3u) == 0, "buffers must be 4-byte aligned");
	if (n == 0) return 0;
	const int vec_ok = (((uintptr_t)d_rgba | (uintptr_t)d_hsva) & 15u) == 0;
	const int grid = grid_for(ctx, (n / 4 + kThreads) / kThreads, 8);
	rgba8_to_hsv8_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
	    reinterpret_cast<const uint32_t *>(d_rgba), n, reinterpret_cast<uint32_t *>(d_hsva), vec_ok);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_feature_norm2_max_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1,
                                        const float *d_f2, int64_t n, double *d_out, void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_out, "null pointer");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), (cudaStream_t)stream));
	if (n == 0) return 0;
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	norm2_max_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_f0, d_f1, d_f2, n,
	                                                              reinterpret_cast<unsigned long long *>(d_out));
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_assign_remap_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int space,
                                     const double *d_lut256, const double *d_centers,
                                     const uint8_t *d_palette_rgb, int K, int preserve_alpha,
                                     uint8_t *d_rgba_out, uint8_t *d_labels, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_centers && d_palette_rgb && d_rgba_out, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(space >= 0 && space <= 2, "unknown colour space");
	CS_REQUIRE(space != CS_SPACE_LAB || d_lut256, "LAB needs the linearisation table");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(((uintptr_t)d_rgba & 3u) == 0 && ((uintptr_t)d_rgba_out & 3u) == 0, "buffers must be 4-byte aligned");
	if (n == 0) return 0;
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	const uint32_t *in = reinterpret_cast<const uint32_t *>(d_rgba);
	uint32_t *out = reinterpret_cast<uint32_t *>(d_rgba_out);
	cudaStream_t st = (cudaStream_t)stream;
	if (space == CS_SPACE_RGB)
		assign_remap_kernel<0><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	else if (space == CS_SPACE_LAB)
		assign_remap_kernel<1><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	else
		assign_remap_kernel<2><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_remap_labels_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, const uint8_t *d_labels,
                                     int64_t n, const uint8_t *d_selpx, int mask_mode, int min_bright,
                                     const uint8_t *d_palette_rgb, int K, int preserve_alpha,
                                     uint8_t *d_rgba_out, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_labels && d_palette_rgb && d_rgba_out, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	if (n == 0) return 0;
	const int grid = grid_for(ctx, (n + kThreads - 1) / kThreads, 8);
	remap_labels_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(
	    reinterpret_cast<const uint32_t *>(d_rgba), d_labels, n, reinterpret_cast<const uint32_t *>(d_selpx), mask_mode,
	    min_bright, d_palette_rgb, K, preserve_alpha, reinterpret_cast<uint32_t *>(d_rgba_out));
	CS_CUDA(cudaGetLastError());
	return 0;
}
