// lab.cu — K1 (sRGB u8 -> CIELAB), K9 (RGB -> OpenCV 8-bit HSV) and K4 (nearest centre +
// palette remap, fused with the colour-space conversion), sm_100a.
//
// K1 replaces skimage.color.rgb2lab (app/processing/color_simplify.py:470, 540, 658, 688, 757,
// 1090-1091); K9 replaces cv2.cvtColor(COLOR_RGB2HSV) on uint8 (:947, 1097-1098); K4 replaces
// sklearn pairwise_distances_argmin_min + the gather/alpha/dstack epilogue (:543-557, 691-705,
// 1106-1121).  All three are streaming kernels: 16-byte loads of 4 RGBA8 pixels per thread,
// conversion in registers, 16-byte stores.
#include "cs_common.cuh"
#include <cstdlib>

namespace cs {
namespace {

constexpr int kThreads = 256;
// below these sizes the table build (~30 us) + its 128 KB load per CTA cost more than the direct kernel's K
// distances per pixel save (direct: ~0.015 ms/MP at K = 16, 0.047 at K = 64; three-phase tiles: ~0.008 / 0.013)
constexpr long long kRemapGridMinPixels = 1 << 21;       // K >= 32
constexpr long long kRemapGridMinPixelsSmallK = 1 << 22;  // K < 32

// skimage.color.colorconv.xyz_from_rgb and the D65 / 2 degree white point
__constant__ double kM[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160,
                             0.072169, 0.019334, 0.119193, 0.950227};
__constant__ double kWhite[3] = {0.95047, 1.0, 1.08883};
__constant__ double kWhiteInv[3] = {1.0 / 0.95047, 1.0, 1.0 / 1.08883};

// cube root to < 1e-15 relative with FOUR special-function / conversion instructions (the unit that bounded K1:
// profiles/r2_ncu_secondary_kernels.md, XU pipe 73 %): z0 = t^(-1/3) from lg2 / ex2 in fp32 (~1e-6), two
// division-free Newton steps for the inverse cube root in fp64, z <- z + z (1 - t z^3) / 3 (quadratic: 1e-6 ->
// 2e-12 -> 1e-23), then cbrt(t) = t z^2.
__device__ __forceinline__ double cbrt_fast(double t) {
	float zf;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(zf) : "f"(__log2f((float)t) * (-1.0f / 3.0f)));
	double z = (double)zf;
#pragma unroll
	for (int it = 0; it < 2; ++it) {
		const double e = fma(-t, z * z * z, 1.0);
		z = fma(z * (1.0 / 3.0), e, z);
	}
	return t * z * z;
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

// ================= K4, grid-filtered (large images; RGB and LAB metrics) =================
// Every input pixel is one of 2^24 colours and the palette is fixed for the whole call, so the decision
// "which centre is nearest" is tabulated per cell of a 32 x 32 x 32 grid over the RGB cube (8 x 8 x 8 colours
// per cell): remap_grid_build_kernel lists the centres that can be nearest for SOME colour of the cell (the
// dominance filter of the Lloyd grid path, csrc/lloyd.cu; for LAB the cell's image is bounded through the
// three monotone functions fx, fy, fz of skimage's xyz2lab, in which the squared distance difference of two
// centres is linear).  remap_grid_kernel then works tile by tile in three phases:
//   1. every pixel: cell lookup; a cell with ONE candidate gives the label (no conversion, no distances);
//      the others are queued in shared memory (warp-aggregated append);
//   2. the queue, densely (all lanes busy although the mixed pixels are scattered): fp32 features, the <= 4
//      candidates' distances, and — when the two best are closer than the bound on the fp32 error — the
//      exact evaluation of the old kernel (fp64 features, all K centres, strict <);
//   3. the finished tile leaves shared memory with 16-byte stores.
// Labels equal assign_remap_kernel's (the fp64 first minimum) for every input.
constexpr int kRgThreads = 512;
constexpr int kRgCells = 32 * 32 * 32;

constexpr int kRgWarps = kRgThreads / 32;
constexpr int kRgSub = 256;  // pixels per warp and step: 8 per lane as two 16-byte groups

struct RgWarp {  // one per warp: nothing in the main loop needs a block-wide barrier
	uint2 queue[kRgSub];     // {pixel, position} of the pixels of mixed cells
	uint32_t outt[kRgSub];   // the finished sub-tile
	uint8_t queue2[kRgSub];  // indices into `queue` of the pixels that need the exact (fp64, all K) evaluation
	uint8_t labt[kRgSub];
};

struct RgConsts {  // what the evaluation of one colour needs (rg_mixed_label, rg_exact_label)
	float4 cf[CS_MAX_K];
	double c64[CS_MAX_K * 3];
	float lutf[256];
	double lutd[256];
};

struct RgSmem {
	uint32_t tab[kRgCells];
	RgWarp w[kRgWarps];
	RgConsts k;
	uint32_t pal[CS_MAX_K];
};

__device__ __forceinline__ void rg_fill_consts(RgConsts &C, const double *__restrict__ lut_g, const double *__restrict__ centers,
                                               int K, bool lab, int tid, int nthreads) {
	for (int i = tid; i < 256; i += nthreads) {
		const double v = lab ? lut_g[i] : 0.0;
		C.lutd[i] = v; C.lutf[i] = (float)v;
	}
	for (int i = tid; i < CS_MAX_K; i += nthreads) {
		const bool ok = i < K;
		const double cx = ok ? centers[3 * i] : 0.0, cy = ok ? centers[3 * i + 1] : 0.0, cz = ok ? centers[3 * i + 2] : 0.0;
		C.c64[3 * i] = cx; C.c64[3 * i + 1] = cy; C.c64[3 * i + 2] = cz;
		C.cf[i] = make_float4((float)cx, (float)cy, (float)cz, 0.f);
	}
}

// skimage's xyz2lab f(): monotone increasing
__device__ __forceinline__ double lab_f64(double v) { return v > 0.008856 ? cbrt(v) : 7.787 * v + 16.0 / 116.0; }

// fp32 CIELAB of one pixel: |L err| <= 1e-4, |a err| <= 6e-4, |b err| <= 3e-4 against the fp64 evaluation
// (fp32 table, 9 FMAs, cbrtf <= 1 ulp; generous by a factor of ~4 — DESIGN.md K4)
// rows of xyz_from_rgb divided by the white point, as compile-time fp32 constants (kM / kWhiteInv live in
// constant memory: using them in the fp32 path would cost an fp64 multiply and a conversion per coefficient)
#define CS_KMW(i) ((i) == 0 ? (float)(0.412453 / 0.95047) : (i) == 1 ? (float)(0.357580 / 0.95047) : (i) == 2 ? (float)(0.180423 / 0.95047) : \
                   (i) == 3 ? 0.212671f : (i) == 4 ? 0.715160f : (i) == 5 ? 0.072169f : \
                   (i) == 6 ? (float)(0.019334 / 1.08883) : (i) == 7 ? (float)(0.119193 / 1.08883) : (float)(0.950227 / 1.08883))

// second half of rgb_to_lab_f32: from the three white-normalised tristimulus values
__device__ __forceinline__ void xyzn_to_lab_f32(const float (&tn)[3], float &L, float &A, float &B) {
	float f[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		const float t = tn[i];
		// cube root of t in (0.008856, 1.1]: 2^(log2(t)/3) from the special-function unit (relative error ~1e-6),
		// then one Newton step y - (y^3 - t) / (3 y^2), which squares that error away (<= 2e-7 after rounding)
		float y0;
		asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__log2f(t) * (1.0f / 3.0f)));
		const float y2 = y0 * y0;
		float r3;  // the correction is ~1e-6 y: an approximate reciprocal (2^-23) changes nothing at fp32
		asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r3) : "f"(3.0f * y2));
		y0 = fmaf(-fmaf(y2, y0, -t), r3, y0);
		f[i] = t > 0.008856f ? y0 : fmaf(7.787f, t, 16.0f / 116.0f);
	}
	L = fmaf(116.f, f[1], -16.f); A = 500.f * (f[0] - f[1]); B = 200.f * (f[1] - f[2]);
}

__device__ __forceinline__ void rgb_to_lab_f32(const float *lutf, uint32_t r, uint32_t g, uint32_t b, float &L, float &A, float &B) {
	const float lr = lutf[r], lg = lutf[g], lb = lutf[b];
	float tn[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) tn[i] = fmaf(lr, CS_KMW(3 * i), fmaf(lg, CS_KMW(3 * i + 1), lb * CS_KMW(3 * i + 2)));
	xyzn_to_lab_f32(tn, L, A, B);
}

template <int SPACE>
__device__ __noinline__ int rg_exact_label(const RgConsts &S, uint32_t w, int K) {
	const int r = w & 0xFF, g = (w >> 8) & 0xFF, b = (w >> 16) & 0xFF;
	double x, y, z;
	if (SPACE == 0) { x = r; y = g; z = b; }
	else rgb_to_lab_f64<false>(S.lutd, r, g, b, x, y, z);
	return remap_exact_label(x, y, z, S.c64, K);
}

// label of a pixel of a mixed cell (phase 2), or -1: needs the exact evaluation (phase 2b)
// the <= 4 candidates of entry e (not an overflow entry) at the fp32 features (x, y, z): label, or -1 = too close
template <int SPACE>
__device__ __forceinline__ int rg_screen(const RgConsts &S, float x, float y, float z, uint32_t e) {
	float best = 3.0e38f, sec = 3.0e38f;
	uint32_t bl = e & 0xFFu;
#pragma unroll
	for (int sl = 0; sl < 4; ++sl) {
		const uint32_t l = (e >> (8 * sl)) & 0xFFu;  // ascending, distinct
		const float4 c = S.cf[l];
		const float dx = x - c.x, dy = y - c.y, dz = z - c.z;
		const float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
		if (d < best) { sec = best; best = d; bl = l; } else sec = fminf(sec, d);
	}
	// |d_fp32 - d| <= 2 sqrt(d) |dx| + |dx|^2 + 3e-7 d with |dx| the feature error (LAB: <= 7e-4; RGB: the fp32
	// rounding of the centres, <= 2.6e-5); two distances
	const float tau = SPACE == 0 ? fmaf(1.2e-4f, sqrtf(sec), fmaf(8e-7f, sec, 1e-6f)) : fmaf(3e-3f, sqrtf(sec), fmaf(8e-7f, sec, 2e-5f));
	if (sec - best <= tau) return -1;
	return (int)bl;
}

template <int SPACE>
__device__ __forceinline__ int rg_mixed_label(const RgConsts &S, uint32_t w, uint32_t e, int K) {
	const uint32_t l0 = e & 0xFFu, l1 = (e >> 8) & 0xFFu;
	if (l0 > l1) return -1;  // more than four candidates
	const uint32_t r = w & 0xFFu, g = (w >> 8) & 0xFFu, b = (w >> 16) & 0xFFu;
	float x, y, z;
	if (SPACE == 0) { x = (float)r; y = (float)g; z = (float)b; }
	else rgb_to_lab_f32(S.lutf, r, g, b, x, y, z);
	return rg_screen<SPACE>(S, x, y, z, e);
}

template <int SPACE>
__global__ void __launch_bounds__(kRgThreads, 1) remap_grid_kernel(
    const uint32_t *__restrict__ rgba, long long n, const double *__restrict__ lut_g, const double *__restrict__ centers,
    const uint8_t *__restrict__ palette, int K, int preserve_alpha, uint32_t *__restrict__ out, uint8_t *__restrict__ labels,
    const uint32_t *__restrict__ table) {
	extern __shared__ __align__(16) unsigned char rg_raw[];
	RgSmem &S = *reinterpret_cast<RgSmem *>(rg_raw);
	const int tid = threadIdx.x, lane = tid & 31;
	for (int i = tid; i < kRgCells / 4; i += kRgThreads)
		reinterpret_cast<uint4 *>(S.tab)[i] = reinterpret_cast<const uint4 *>(table)[i];
	rg_fill_consts(S.k, lut_g, centers, K, SPACE == 1, tid, kRgThreads);
	for (int i = tid; i < CS_MAX_K; i += kRgThreads)
		S.pal[i] = i < K ? ((uint32_t)palette[3 * i] | ((uint32_t)palette[3 * i + 1] << 8) | ((uint32_t)palette[3 * i + 2] << 16)) : 0u;
	__syncthreads();
	// ---- main loop: every warp on its own 256-pixel sub-tiles, three phases separated by warp barriers only ----
	RgWarp &W = S.w[tid >> 5];
	const long long nsub = (n + kRgSub - 1) / kRgSub;
	const long long wstride = (long long)gridDim.x * kRgWarps;
	const uint32_t lt = (1u << lane) - 1u;
	uint4 nxt[2];
	auto fetch = [&](long long sub) {
		const long long base = sub * kRgSub;
#pragma unroll
		for (int u = 0; u < 2; ++u) {
			const long long p0 = base + (u * 32 + lane) * 4;
			if (p0 + 4 <= n) nxt[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba + p0));
			else {
				uint32_t t4[4];
#pragma unroll
				for (int q = 0; q < 4; ++q) t4[q] = p0 + q < n ? rgba[p0 + q] : 0u;  // alpha 0: a no-op pixel
				nxt[u] = make_uint4(t4[0], t4[1], t4[2], t4[3]);
			}
		}
	};
	long long sub = (long long)blockIdx.x * kRgWarps + (tid >> 5);
	if (sub < nsub) fetch(sub);
	for (; sub < nsub; sub += wstride) {
		const long long base = sub * kRgSub;
		// ---- phase 1: classify; pure cells and transparent pixels are final, the others are queued ----
		int cnt = 0;
#pragma unroll
		for (int u = 0; u < 2; ++u) {
			const int p0 = (u * 32 + lane) * 4;
			const uint32_t w4[4] = {nxt[u].x, nxt[u].y, nxt[u].z, nxt[u].w};
			uint32_t o4[4], lab4 = 0u;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t w = w4[q], a = w >> 24;
				const uint32_t a_out = (preserve_alpha ? a : (a > 128u ? 255u : 0u)) << 24;
				const uint32_t cell = ((w >> 3) & 0x1Fu) | ((w >> 6) & 0x3E0u) | ((w >> 9) & 0x7C00u);
				const uint32_t e = S.tab[cell];
				const bool opaque = a > 0u;
				const bool mixed = opaque && e != __byte_perm(e, 0u, 0x0000);
				o4[q] = (opaque ? S.pal[e & 0xFFu] : 0u) | a_out;
				lab4 |= (opaque ? (e & 0xFFu) : 255u) << (8 * q);
				const uint32_t m = __ballot_sync(0xffffffffu, mixed);
				if (mixed) W.queue[cnt + __popc(m & lt)] = make_uint2(w, (uint32_t)(p0 + q));
				cnt += __popc(m);
			}
			*reinterpret_cast<uint4 *>(&W.outt[p0]) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
			*reinterpret_cast<uint32_t *>(&W.labt[p0]) = lab4;
		}
		// the next sub-tile's pixels travel while this one is finished
		if (sub + wstride < nsub) fetch(sub + wstride);
		__syncwarp();
		// ---- phase 2: the mixed pixels, densely ----
		int cnt2 = 0;
		for (int i0 = 0; i0 < cnt; i0 += 32) {
			const int i = i0 + lane;
			int l = 0;
			uint2 qe = make_uint2(0u, 0u);
			if (i < cnt) {
				qe = W.queue[i];
				const uint32_t w = qe.x;
				const uint32_t cell = ((w >> 3) & 0x1Fu) | ((w >> 6) & 0x3E0u) | ((w >> 9) & 0x7C00u);
				l = rg_mixed_label<SPACE>(S.k, w, S.tab[cell], K);
				if (l >= 0) {
					W.outt[qe.y] = S.pal[l] | (W.outt[qe.y] & 0xFF000000u);
					W.labt[qe.y] = (uint8_t)l;
				}
			}
			const uint32_t m = __ballot_sync(0xffffffffu, l < 0);
			if (l < 0) W.queue2[cnt2 + __popc(m & lt)] = (uint8_t)i;
			cnt2 += __popc(m);
		}
		__syncwarp();
		// ---- phase 2b: the few pixels the fp32 screen could not decide, again densely ----
		for (int i = lane; i < cnt2; i += 32) {
			const uint2 qe = W.queue[W.queue2[i]];
			const int l = rg_exact_label<SPACE>(S.k, qe.x, K);
			W.outt[qe.y] = S.pal[l] | (W.outt[qe.y] & 0xFF000000u);
			W.labt[qe.y] = (uint8_t)l;
		}
		__syncwarp();
		// ---- phase 3: the sub-tile leaves with 16-byte stores ----
#pragma unroll
		for (int u = 0; u < 2; ++u) {
			const int p0 = (u * 32 + lane) * 4;
			const long long g0 = base + p0;
			if (g0 + 4 <= n) stg_stream_u4(reinterpret_cast<uint4 *>(out + g0), *reinterpret_cast<const uint4 *>(&W.outt[p0]));
			else
				for (int q = 0; q < 4; ++q)
					if (g0 + q < n) out[g0 + q] = W.outt[p0 + q];
		}
		if (labels && lane < 16) {
			const int p0 = lane * 16;
			const long long g0 = base + p0;
			if (g0 + 16 <= n) *reinterpret_cast<uint4 *>(labels + g0) = *reinterpret_cast<const uint4 *>(&W.labt[p0]);
			else
				for (int q = 0; q < 16; ++q)
					if (g0 + q < n) labels[g0 + q] = W.labt[p0 + q];
		}
		__syncwarp();
	}
}

// One warp per cell.  Features phi of a colour: (r, g, b) for RGB, (fx, fy, fz) for LAB; both monotone in each
// of r, g, b, so the cell's image lies in the box [phi(low corner), phi(high corner)].  With lab = A phi + a0
// (RGB: identity) the difference d_w - d_k = |c_w|^2 - |c_k|^2 + 2 a0.u + 2 phi.(A^T u), u = c_k - c_w, is
// linear in phi: k is dominated by w when its maximum over the box is negative.
// G lanes per cell: a half-warp when K <= 16 (two cells per warp — every step below is per group), else a warp.
template <int SPACE, int G>
__global__ void __launch_bounds__(256) remap_grid_build_kernel(const double *__restrict__ centers, int K,
                                                               const double *__restrict__ lut_g, uint32_t *__restrict__ table) {
	__shared__ double c[CS_MAX_K * 3], qn[CS_MAX_K], lut[256];
	__shared__ int clist[8][32];
	for (int i = threadIdx.x; i < K * 3; i += 256) c[i] = centers[i];
	for (int i = threadIdx.x; i < 256; i += 256) lut[i] = SPACE == 1 ? lut_g[i] : 0.0;
	__syncthreads();
	for (int i = threadIdx.x; i < K; i += 256) qn[i] = c[3 * i] * c[3 * i] + c[3 * i + 1] * c[3 * i + 1] + c[3 * i + 2] * c[3 * i + 2];
	__syncthreads();
	constexpr int kPerWarp = 32 / G;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gl = lane & (G - 1), grp = lane / G;
	const uint32_t below = (1u << gl) - 1u;
	auto group_bits = [&](uint32_t m) { return G == 32 ? m : (m >> (G * grp)) & ((1u << (G & 31)) - 1u); };
	int *mine = clist[wib] + grp * G;
	auto alpha = [&](const double (&u)[3], double (&al)[3], double &a0u) {
		if (SPACE == 1) {
			al[0] = 500.0 * u[1]; al[1] = 116.0 * u[0] - 500.0 * u[1] + 200.0 * u[2]; al[2] = -200.0 * u[2];
			a0u = -16.0 * u[0];
		} else { al[0] = u[0]; al[1] = u[1]; al[2] = u[2]; a0u = 0.0; }
	};
	// kRgCells is a multiple of every step: the groups of a warp run the same number of iterations
	for (int cell = (blockIdx.x * 8 + wib) * kPerWarp + grp; cell < kRgCells; cell += gridDim.x * 8 * kPerWarp) {
		const int r0 = (cell & 31) * 8, g0 = ((cell >> 5) & 31) * 8, b0 = (cell >> 10) * 8;
		double lo[3], hi[3];
		if (SPACE == 1) {
			// group lanes 0-2: (fx, fy, fz) of the low corner, 3-5: of the high corner — one cube root per lane
			const int cmp = gl % 3, up = (gl / 3) & 1 ? 7 : 0;
			const double v = lab_f64((lut[r0 + up] * kM[3 * cmp] + lut[g0 + up] * kM[3 * cmp + 1] + lut[b0 + up] * kM[3 * cmp + 2]) / kWhite[cmp]);
#pragma unroll
			for (int j = 0; j < 3; ++j) { lo[j] = __shfl_sync(0xffffffffu, v, grp * G + j); hi[j] = __shfl_sync(0xffffffffu, v, grp * G + 3 + j); }
		} else { lo[0] = r0; lo[1] = g0; lo[2] = b0; hi[0] = r0 + 7; hi[1] = g0 + 7; hi[2] = b0 + 7; }
#pragma unroll
		for (int j = 0; j < 3; ++j) { const double pad = 1e-9 * (1.0 + fabs(hi[j])); lo[j] -= pad; hi[j] += pad; }
		auto dominated = [&](int k, int w) {  // w closer than k everywhere in the box
			const double u[3] = {c[3 * k] - c[3 * w], c[3 * k + 1] - c[3 * w + 1], c[3 * k + 2] - c[3 * w + 2]};
			double al[3], a0u;
			alpha(u, al, a0u);
			double m = qn[w] - qn[k] + 2.0 * a0u;
#pragma unroll
			for (int j = 0; j < 3; ++j) m += 2.0 * fmax(al[j] * lo[j], al[j] * hi[j]);
			return m < -1e-7 * (1.0 + qn[w] + qn[k]);
		};
		// w* = centre nearest to the middle of the box
		double mid[3];
		{
			const double p[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
			if (SPACE == 1) { mid[0] = 116.0 * p[1] - 16.0; mid[1] = 500.0 * (p[0] - p[1]); mid[2] = 200.0 * (p[1] - p[2]); }
			else { mid[0] = p[0]; mid[1] = p[1]; mid[2] = p[2]; }
		}
		double bd = 1e300;
		int bw = 0x7fffffff;
		for (int k = gl; k < K; k += G) {
			const double dx = mid[0] - c[3 * k], dy = mid[1] - c[3 * k + 1], dz = mid[2] - c[3 * k + 2];
			const double d = dx * dx + dy * dy + dz * dz;
			if (d < bd) { bd = d; bw = k; }
		}
#pragma unroll
		for (int o = G / 2; o > 0; o >>= 1) {
			const double od = __shfl_xor_sync(0xffffffffu, bd, o);
			const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
			if (od < bd || (od == bd && ow < bw)) { bd = od; bw = ow; }
		}
		// candidates = not dominated by w* (ascending labels), at most G kept
		int cnt = 0;
		for (int k0 = 0; k0 < K; k0 += G) {
			const int k = k0 + gl;
			const bool cand = k < K && (k == bw || !dominated(k, bw));
			const uint32_t m = group_bits(__ballot_sync(0xffffffffu, cand));
			if (cand) {
				const int pos = cnt + __popc(m & below);
				if (pos < G) mine[pos] = k;
			}
			cnt += __popc(m);
		}
		__syncwarp();
		// refine: drop a candidate that another candidate dominates (w* alone is a weak filter near a face)
		{
			const bool refine = cnt > 1 && cnt <= G;
			bool keep = gl < cnt;
			if (refine && keep)
				for (int j = 0; j < cnt; ++j)
					if (j != gl && dominated(mine[gl], mine[j])) { keep = false; break; }
			const uint32_t m = group_bits(__ballot_sync(0xffffffffu, keep));
			const int mylab = gl < cnt && cnt <= G ? mine[gl] : 0;
			__syncwarp();
			if (refine) {
				if (keep) mine[__popc(m & below)] = mylab;
				cnt = __popc(m);
			}
			__syncwarp();
		}
		if (gl == 0) {
			uint32_t entry;
			if (cnt == 1) {
				entry = (uint32_t)mine[0] * 0x01010101u;
			} else if (cnt <= 4 && K >= 4) {
				int l4[4], m4 = cnt;
				for (int i = 0; i < cnt; ++i) l4[i] = mine[i];
				for (int k = 0; m4 < 4 && k < K; ++k) {  // pad with distinct real centres
					bool in = false;
					for (int i = 0; i < m4; ++i) in = in || l4[i] == k;
					if (!in) l4[m4++] = k;
				}
				for (int i = 1; i < 4; ++i)  // ascending
					for (int j = i; j > 0 && l4[j] < l4[j - 1]; --j) { const int t = l4[j]; l4[j] = l4[j - 1]; l4[j - 1] = t; }
				entry = (uint32_t)l4[0] | ((uint32_t)l4[1] << 8) | ((uint32_t)l4[2] << 16) | ((uint32_t)l4[3] << 24);
			} else {
				entry = 0x00000001u;  // byte0 > byte1: evaluate all K
			}
			table[cell] = entry;
		}
		__syncwarp();
	}
}

// ================= K4, colour table (images >= 16 MP) =================
// From 2^24 pixels up there are at least as many pixels as colours, so the mixed cells are decided once per
// COLOUR instead of once per pixel: remap_lut_build_kernel evaluates the 512 colours of every mixed cell (the
// same fp32 screen + exact fp64 evaluation as phase 2 / 2b above, so the labels are the direct kernel's) into
// a byte table indexed by the colour itself (b << 16 | g << 8 | r: the low three bytes of the pixel) — 16 MB, only
// the mixed cells' colours are ever written or read, and they stay in L2.  The same kernel, knowing every label of the cell,
// writes the table the pass keeps in shared memory: one byte per FINE cell of 8 x 4 x 4 colours (32 x 64 x 64
// cells, 128 KB) — the label when all 128 colours agree, 255 = look the colour up.  This table is exact, where the
// dominance filter is only safe: on a random 16-colour palette the filter calls 52 % of the 8 x 8 x 8 cells mixed,
// 22 % of them are, and 12 % of the fine cells.  remap_lut_kernel is then a pure streaming pass: 16-byte pixel
// loads, one shared-memory byte per pixel, one byte gather from L2 for the pixels of mixed fine cells (each a
// 32-byte sector through the L1 tag stage, which is what bounds the pass), 16-byte stores — no colour
// conversion, no distances, no queues.
constexpr long long kRemapLutMinPixels = 1 << 24;
constexpr int kRlThreads = 1024;
constexpr int kRlFine = 32 * 64 * 64;  // [b >> 2][g >> 2][r >> 3]

struct RlSmem {
	uint8_t fine[kRlFine];
	uint32_t pal[CS_MAX_K];
};

template <int SPACE>
__global__ void __launch_bounds__(256) remap_lut_build_kernel(const double *__restrict__ centers, int K,
                                                              const double *__restrict__ lut_g,
                                                              const uint32_t *__restrict__ table, uint8_t *__restrict__ lut8,
                                                              uint8_t *__restrict__ fine) {
	__shared__ RgConsts C;
	rg_fill_consts(C, lut_g, centers, K, SPACE == 1, threadIdx.x, 256);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	for (int cell = blockIdx.x * 8 + wib; cell < kRgCells; cell += gridDim.x * 8) {
		const uint32_t e = table[cell];
		// the four fine cells of this cell: (b & 4, g & 4) = (bh, gh)
		auto fine_idx = [&](int bh, int gh) { return ((2 * (cell >> 10) + bh) << 11) | ((2 * ((cell >> 5) & 31) + gh) << 5) | (cell & 31); };
		if (e == __byte_perm(e, 0u, 0x0000)) {  // one candidate: the pass never looks this row up ...
			if ((e & 0xFFu) == 255u) {          // ... unless the label is the "look it up" byte itself (K = 256)
				const uint32_t c0 = ((uint32_t)(cell & 31) << 3) | ((uint32_t)(((cell >> 5) & 31) * 8 + 2 * (lane & 3)) << 8) |
				                    ((uint32_t)((cell >> 10) * 8 + (lane >> 2)) << 16);
				*reinterpret_cast<uint2 *>(lut8 + c0) = make_uint2(~0u, ~0u);
				*reinterpret_cast<uint2 *>(lut8 + c0 + 256) = make_uint2(~0u, ~0u);
			}
			if (lane < 4) fine[fine_idx(lane >> 1, lane & 1)] = (uint8_t)(e & 0xFFu);
			continue;
		}
		const uint32_t base = ((uint32_t)(cell & 31) << 3) | ((uint32_t)((cell >> 5) & 31) << 11) | ((uint32_t)(cell >> 10) << 19);
		// lane: the 16 colours (b & 7) = lane >> 2, (g & 7) = 2 (lane & 3) + {0, 1}, (r & 7) = 0..7 — two 8-byte stores
		const uint32_t wl = base | ((uint32_t)(2 * (lane & 3)) << 8) | ((uint32_t)(lane >> 2) << 16);  // + r & 7, + (g & 1) << 8
		uint32_t pk[4] = {0u, 0u, 0u, 0u};
		uint32_t hard = (e & 0xFFu) > ((e >> 8) & 0xFFu) ? 0xFFFFu : 0u;  // more than four candidates: all 16 exactly
		if (!hard) {
			// fp32 screen, fully unrolled: the green and blue terms of the tristimulus sums are shared by 8 colours
			float part[2][3], lr[8];
			if (SPACE == 1) {
				const float lb = C.lutf[(wl >> 16) & 0xFFu];
#pragma unroll
				for (int gs = 0; gs < 2; ++gs) {
					const float lg = C.lutf[((wl >> 8) & 0xFFu) + gs];
#pragma unroll
					for (int i = 0; i < 3; ++i) part[gs][i] = fmaf(lg, CS_KMW(3 * i + 1), lb * CS_KMW(3 * i + 2));
				}
#pragma unroll
				for (int r = 0; r < 8; ++r) lr[r] = C.lutf[(wl & 0xFFu) + r];
			}
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				float x, y, z;
				if (SPACE == 1) {
					float tn[3];
#pragma unroll
					for (int i = 0; i < 3; ++i) tn[i] = fmaf(lr[j & 7], CS_KMW(3 * i), part[j >> 3][i]);
					xyzn_to_lab_f32(tn, x, y, z);
				} else {
					x = (float)((wl & 0xFFu) + (j & 7)); y = (float)(((wl >> 8) & 0xFFu) + (j >> 3)); z = (float)((wl >> 16) & 0xFFu);
				}
				const int l = rg_screen<SPACE>(C, x, y, z, e);
				if (l < 0) hard |= 1u << j;
				else pk[j >> 2] |= (uint32_t)l << (8 * (j & 3));
			}
		}
		while (hard) {  // near ties of the screen (rare), or the whole row of an overflow cell
			const int j = __ffs(hard) - 1;
			hard &= hard - 1u;
			const int l = rg_exact_label<SPACE>(C, wl + (uint32_t)(j & 7) + ((uint32_t)(j >> 3) << 8), K);
			pk[j >> 2] |= (uint32_t)l << (8 * (j & 3));
		}
		*reinterpret_cast<uint2 *>(lut8 + wl) = make_uint2(pk[0], pk[1]);        // wl = the colour of j = 0
		*reinterpret_cast<uint2 *>(lut8 + wl + 256) = make_uint2(pk[2], pk[3]);  // g + 1
		// fine cells: the lane's 16 colours lie in fine cell (bh, gh) = (lane >> 4, (lane >> 1) & 1); the 8 lanes of a
		// fine cell differ in lane bits 0, 2, 3
		const uint32_t first = pk[0] & 0xFFu, rep4 = first * 0x01010101u;
		uint32_t v = (pk[0] == rep4 && pk[1] == rep4 && pk[2] == rep4 && pk[3] == rep4 && first != 255u) ? first : 0x100u;
#pragma unroll
		for (int m = 1; m <= 8; m = m == 1 ? 4 : m * 2) {
			const uint32_t o = __shfl_xor_sync(0xffffffffu, v, m);
			v = o == v ? v : 0x100u;
		}
		if ((lane & 0x0D) == 0) fine[fine_idx(lane >> 4, (lane >> 1) & 1)] = (uint8_t)(v > 255u ? 255u : v);
	}
}

__device__ __forceinline__ uint32_t ldg_nc_u8(const uint8_t *p) {
	uint32_t v;
	asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

template <bool PA>  // preserve_alpha
__global__ void __launch_bounds__(kRlThreads, 1) remap_lut_kernel(
    const uint32_t *__restrict__ rgba, long long n, const uint8_t *__restrict__ palette, int K,
    uint32_t *__restrict__ out, uint8_t *__restrict__ labels, const uint8_t *__restrict__ fine,
    const uint8_t *__restrict__ lut8) {
	extern __shared__ __align__(16) unsigned char rl_raw[];
	RlSmem &S = *reinterpret_cast<RlSmem *>(rl_raw);
	const int tid = threadIdx.x;
	for (int i = tid; i < kRlFine / 16; i += kRlThreads)
		reinterpret_cast<uint4 *>(S.fine)[i] = reinterpret_cast<const uint4 *>(fine)[i];
	for (int i = tid; i < CS_MAX_K; i += kRlThreads)
		S.pal[i] = i < K ? ((uint32_t)palette[3 * i] | ((uint32_t)palette[3 * i + 1] << 8) | ((uint32_t)palette[3 * i + 2] << 16)) : 0u;
	__syncthreads();
	const long long ngroups = (n + 3) / 4;  // 4 pixels = one 16-byte group per thread and step
	const long long stride = (long long)gridDim.x * kRlThreads;
	auto one_group = [&](long long gidx, const uint4 px) {
		const uint32_t w4[4] = {px.x, px.y, px.z, px.w};
		uint32_t lab[4], o4[4], l4 = 0u;
#pragma unroll
		for (int q = 0; q < 4; ++q) {  // the (up to four) gathers of the group leave together
			const uint32_t w = w4[q];
			lab[q] = S.fine[((w >> 7) & 0x1F800u) | ((w >> 5) & 0x7E0u) | ((w >> 3) & 0x1Fu)];
			if (w > 0x00FFFFFFu && lab[q] == 255u) lab[q] = ldg_nc_u8(lut8 + (w & 0x00FFFFFFu));
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const uint32_t w = w4[q];
			const uint32_t a_out = PA ? (w & 0xFF000000u) : (w > 0x80FFFFFFu ? 0xFF000000u : 0u);  // alpha, or (alpha > 128) * 255
			const bool opaque = w > 0x00FFFFFFu;
			o4[q] = (opaque ? S.pal[lab[q]] : 0u) | a_out;
			l4 |= (opaque ? lab[q] : 255u) << (8 * q);
		}
		const long long p0 = gidx * 4;
		if (p0 + 4 <= n) {
			stg_stream_u4(reinterpret_cast<uint4 *>(out + p0), make_uint4(o4[0], o4[1], o4[2], o4[3]));
			if (labels) *reinterpret_cast<uint32_t *>(labels + p0) = l4;
		} else {
			for (int q = 0; q < 4; ++q)
				if (p0 + q < n) { out[p0 + q] = o4[q]; if (labels) labels[p0 + q] = (uint8_t)(l4 >> (8 * q)); }
		}
	};
	auto fetch = [&](long long gidx) {
		const long long p0 = gidx * 4;
		if (p0 + 4 <= n) return ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba + p0));
		uint32_t t4[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) t4[q] = p0 + q < n ? rgba[p0 + q] : 0u;
		return make_uint4(t4[0], t4[1], t4[2], t4[3]);
	};
	// Main loop: kRlDepth complete groups per thread and step — the 16-byte loads leave together (64 KB in flight
	// per SM), then all the table bytes and the gathers they ask for, then the outputs.
	constexpr int kRlDepth = 4;
	const long long nwhole = n / 4;
	long long g0 = (long long)blockIdx.x * kRlThreads + tid;
	for (; g0 + (kRlDepth - 1) * stride < nwhole; g0 += kRlDepth * stride) {
		uint4 px[kRlDepth];
#pragma unroll
		for (int u = 0; u < kRlDepth; ++u) px[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(rgba + (g0 + u * stride) * 4));
		uint32_t lab[kRlDepth][4];
#pragma unroll
		for (int u = 0; u < kRlDepth; ++u) {
			const uint32_t w4[4] = {px[u].x, px[u].y, px[u].z, px[u].w};
#pragma unroll
			for (int q = 0; q < 4; ++q) lab[u][q] = S.fine[((w4[q] >> 7) & 0x1F800u) | ((w4[q] >> 5) & 0x7E0u) | ((w4[q] >> 3) & 0x1Fu)];
		}
#pragma unroll
		for (int u = 0; u < kRlDepth; ++u) {
			const uint32_t w4[4] = {px[u].x, px[u].y, px[u].z, px[u].w};
#pragma unroll
			for (int q = 0; q < 4; ++q)
				if (w4[q] > 0x00FFFFFFu && lab[u][q] == 255u) lab[u][q] = ldg_nc_u8(lut8 + (w4[q] & 0x00FFFFFFu));
		}
#pragma unroll
		for (int u = 0; u < kRlDepth; ++u) {
			const uint32_t w4[4] = {px[u].x, px[u].y, px[u].z, px[u].w};
			uint32_t o4[4], l4 = 0u;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t w = w4[q];
				const uint32_t a_out = PA ? (w & 0xFF000000u) : (w > 0x80FFFFFFu ? 0xFF000000u : 0u);
				const bool opaque = w > 0x00FFFFFFu;
				o4[q] = (opaque ? S.pal[lab[u][q]] : 0u) | a_out;
				l4 |= (opaque ? lab[u][q] : 255u) << (8 * q);
			}
			const long long p0 = (g0 + u * stride) * 4;
			stg_stream_u4(reinterpret_cast<uint4 *>(out + p0), make_uint4(o4[0], o4[1], o4[2], o4[3]));
			if (labels) *reinterpret_cast<uint32_t *>(labels + p0) = l4;
		}
	}
	for (; g0 < ngroups; g0 += stride) one_group(g0, fetch(g0));  // the last steps, and the ragged last group
}

template <int SPACE>
void launch_remap_grid_build(int grid, cudaStream_t st, const double *d_centers, int K, const double *d_lut256, uint32_t *d_tab) {
	if (K <= 16) remap_grid_build_kernel<SPACE, 16><<<grid, 256, 0, st>>>(d_centers, K, d_lut256, d_tab);
	else remap_grid_build_kernel<SPACE, 32><<<grid, 256, 0, st>>>(d_centers, K, d_lut256, d_tab);
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
	static const bool no_grid = getenv("CS_NO_GRID") != nullptr;  // development switch
	if (!no_grid && ctx->remap_policy >= 0 && (n >= (K >= 32 ? kRemapGridMinPixels : kRemapGridMinPixelsSmallK) || ctx->remap_policy > 0) && K >= 4 && space != CS_SPACE_HSV && (((uintptr_t)d_rgba | (uintptr_t)d_rgba_out) & 15u) == 0 &&
	    (!d_labels || ((uintptr_t)d_labels & 15u) == 0)) {
		// grid-filtered paths: candidate table over the RGB cube (128 KB, built per call), then either the
		// three-phase tiles or (>= 16 MP) the per-colour table of the mixed cells and a streaming pass
		if (!ctx->d_remap_tab) CS_CUDA(cudaMalloc(&ctx->d_remap_tab, sizeof(uint32_t) * kRgCells));
		const int bgrid = grid_for(ctx, kRgCells / 8, 4);
		if (ctx->remap_policy == 2 || (ctx->remap_policy == 0 && n >= kRemapLutMinPixels)) {
			if (!ctx->d_remap_lut) CS_CUDA(cudaMalloc(&ctx->d_remap_lut, (size_t)kRgCells * 512 + kRlFine));
			uint8_t *d_fine = ctx->d_remap_lut + (size_t)kRgCells * 512;
			static bool attr_l = false;
			if (!attr_l) {
				CS_CUDA(cudaFuncSetAttribute(remap_lut_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RlSmem)));
				CS_CUDA(cudaFuncSetAttribute(remap_lut_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RlSmem)));
				attr_l = true;
			}
			const long long nblk = (n / 4 + kRlThreads - 1) / kRlThreads;
			const int pgrid = (int)(nblk < ctx->sm_count ? (nblk < 1 ? 1 : nblk) : ctx->sm_count);
			const int lgrid = grid_for(ctx, kRgCells / 8, 8);
			if (space == CS_SPACE_RGB) {
				launch_remap_grid_build<0>(bgrid, st, d_centers, K, d_lut256, ctx->d_remap_tab);
				remap_lut_build_kernel<0><<<lgrid, 256, 0, st>>>(d_centers, K, d_lut256, ctx->d_remap_tab, ctx->d_remap_lut, d_fine);
			} else {
				launch_remap_grid_build<1>(bgrid, st, d_centers, K, d_lut256, ctx->d_remap_tab);
				remap_lut_build_kernel<1><<<lgrid, 256, 0, st>>>(d_centers, K, d_lut256, ctx->d_remap_tab, ctx->d_remap_lut, d_fine);
			}
			if (preserve_alpha)
				remap_lut_kernel<true><<<pgrid, kRlThreads, sizeof(RlSmem), st>>>(in, n, d_palette_rgb, K, out, d_labels, d_fine, ctx->d_remap_lut);
			else
				remap_lut_kernel<false><<<pgrid, kRlThreads, sizeof(RlSmem), st>>>(in, n, d_palette_rgb, K, out, d_labels, d_fine, ctx->d_remap_lut);
			CS_CUDA(cudaGetLastError());
			return 0;
		}
		const long long ntiles = (n + (long long)kRgSub * kRgWarps - 1) / ((long long)kRgSub * kRgWarps);
		const int kgrid = (int)(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
		if (space == CS_SPACE_RGB) {
			static bool attr0 = false;
			if (!attr0) { CS_CUDA(cudaFuncSetAttribute(remap_grid_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RgSmem))); attr0 = true; }
			launch_remap_grid_build<0>(bgrid, st, d_centers, K, d_lut256, ctx->d_remap_tab);
			remap_grid_kernel<0><<<kgrid, kRgThreads, sizeof(RgSmem), st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels, ctx->d_remap_tab);
		} else {
			static bool attr1 = false;
			if (!attr1) { CS_CUDA(cudaFuncSetAttribute(remap_grid_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RgSmem))); attr1 = true; }
			launch_remap_grid_build<1>(bgrid, st, d_centers, K, d_lut256, ctx->d_remap_tab);
			remap_grid_kernel<1><<<kgrid, kRgThreads, sizeof(RgSmem), st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels, ctx->d_remap_tab);
		}
		CS_CUDA(cudaGetLastError());
		return 0;
	}
	if (space == CS_SPACE_RGB)
		assign_remap_kernel<0><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	else if (space == CS_SPACE_LAB)
		assign_remap_kernel<1><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	else
		assign_remap_kernel<2><<<grid, kThreads, 0, st>>>(in, n, d_lut256, d_centers, d_palette_rgb, K, preserve_alpha, out, d_labels);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_remap_set_policy(cs_ctx *ctx, int policy) {
	CS_REQUIRE(ctx, "null context");
	CS_REQUIRE(policy >= -1 && policy <= 2, "policy must be -1, 0, 1 or 2");
	ctx->remap_policy = policy;
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
