// lloyd.cu — K2/K3: one Lloyd iteration (assign + update) for 3-feature k-means, sm_100a.
//
// Replaces sklearn's lloyd_iter_chunked_dense/_update_chunk_dense
// (sklearn/cluster/_k_means_lloyd.pyx:23-218) and the M-step tail in
// sklearn/cluster/_k_means_common.pyx:167-311, which the reference reaches through
// KMeans.fit at app/processing/color_simplify.py:79-80, 669-675, 811-812, 992-993.
//
// Shape of the kernel (one persistent CTA per SM, 16 consumer warps + 1 producer warp):
//   prologue       : barrier set-up, accumulator zeroing and the first ring loads run BEFORE
//                    griddepcontrol.wait, so in a chained launch (CS_LLOYD_CHAINED, programmatic
//                    dependent launch) they overlap the previous iteration's combine + M-step tail;
//   producer lane  : 1-D bulk async copies (cp.async.bulk, the TMA path without a tensor
//                    map) of the next pixel tile of each feature plane into a STAGES-deep
//                    shared-memory ring, completion on an mbarrier per stage;
//   consumer warps : LDS.128 of 4 consecutive pixels per plane, distances to all centres
//                    with packed fp32x2 FMAs (FFMA2: two centres per instruction, the pixel
//                    broadcast), integer keys (bits(v) << b) + k so that the argmin is a chain
//                    of unsigned mins with the label in the low bits (float keys with the index
//                    in the low mantissa bits for K >= 128), label bytes packed with PRMT and
//                    stored as one coalesced 32-bit word per 4 pixels;
//   update         : lane-private {sum0,sum1,sum2,count} float4 slots in shared memory
//                    (conflict-free, no atomics), folded to fp64 once per CTA (one warp per
//                    row of 32 slots), written as a per-CTA partial; the last CTA to finish sums
//                    the partials in block order (deterministic), exchanges them with the other
//                    GPUs of a sharded run through peer-mapped mailboxes (self-validating words),
//                    and, in the fused entry points, runs the M-step tail from shared memory.
// HBM traffic per pixel: 12 B read (3 fp32 planes) + 1 B written (u8 label) = 13 B.
#include "cs_common.cuh"
#include <cstdlib>

namespace cs {
namespace {

constexpr int kAccBytesPerWarp = 8192;
constexpr int kSmemBudget = 232448 - 1024;  // 227 KB opt-in limit minus static shared + slack

enum { FM_F32 = 0, FM_RGBA8 = 1 };

#ifndef CS_GRID16_CAP
#define CS_GRID16_CAP kGridCap  // cells of the K <= 16 table (measured: 11520 cells 0.2758 ms, 9984 0.2712: the build and the table copy grow faster than the crowded cells shrink)
#endif
#ifndef CS_GRID64_COPIES
#define CS_GRID64_COPIES 1  // copies of the centre table in the K = 64 GRID kernel (each copy beyond the first costs 256 cells)
#endif

// Kernel shape: NW consumer warps (+1 producer warp), U groups of 4 pixels per consumer
// thread per tile, centre table in registers (KP <= 16) or broadcast from shared memory.
// PP_ = pixels per pass over the centre table (0 = all 4 U at once), MINM_ = how the two keys of a centre pair
// enter the running minimum: 0 = min(min(best,k0),k1) (ptxas forms 3-input mins), 1 = two separate 2-input mins.
template <int NW_, int U_, bool TABREG_, int PP_ = 0, int MINM_ = 0> struct Var {
	static constexpr int NW = NW_, U = U_;
	static constexpr bool TABREG = TABREG_;
	static constexpr int PP = PP_ == 0 ? 4 * U_ : PP_, MINM = MINM_;
	static constexpr int NC = NW * 32, THREADS = NC + 32, TILE = NC * 4 * U;
};
using VarLargeK = Var<16, 1, false, 0, 1>;  // KP >= 32: 24 KB stages, 3-deep ring beside 128 KB of slots
// KP <= 16: 8 pixels per thread per tile, the centre table walked twice (4 pixels per pass), two separate
// 2-input mins per centre pair — the fastest of the arrangements swept on the B200 (profiles/, sweep notes)
using VarSmallK = Var<16, 2, false, 4, 1>;

template <int KP> struct KCfg {
	static constexpr int kCopies = (kAccBytesPerWarp / (KP * 16)) > 32 ? 32 : (kAccBytesPerWarp / (KP * 16));
	static constexpr int kPhases = 32 / kCopies;
	static constexpr int kBits = KP == 8 ? 3 : KP == 16 ? 4 : KP == 32 ? 5 : KP == 64 ? 6 : KP == 128 ? 7 : 8;
	// GRID kernels with K <= 32 keep 8 copies of the centre table (bank-conflict-free gathers, see grid_key3):
	// 2 / 4 KB.  K <= 16 has no pool (eight nibble candidates per cell entry), K = 32 pays with 384 cells (the
	// table and the pool are sized per KP).  At K = 64 the 8 KB would cost 1792 cells: measured again after the
	// rare-path rewrite, 0.671 / 0.657 / 0.658 ms per 64 MP iteration with 8 / 4 / 1 copies — one copy there.
	static constexpr int kGridPoolUsed = KP <= 16 ? 0 : kGridPool;  // K <= 16: eight candidates per cell entry, no pool — the words go to cells
	static constexpr int kGridTabCopies = KP <= 32 ? 8 : CS_GRID64_COPIES;
	static constexpr int kGridTabShift = kGridTabCopies == 8 ? 7 : kGridTabCopies == 4 ? 6 : kGridTabCopies == 2 ? 5 : 4;  // log2(16 * copies)
	static constexpr int kGridCapUsed = KP == 32 ? 9600 : KP == 64 ? kGridCap - 256 * (CS_GRID64_COPIES - 1) : CS_GRID16_CAP;
};

// Geometry of the cell grid of the grid-filtered assignment (see assign_grid): cell index of a pixel along
// axis j = floor(sat(x_j * s[j] + o[j]) * gs[j]) with sat() the clamp to [0, 1]; linear index = ix + g[0] * (iy + g[1] * iz).
struct GridGeom {
	float s[3], o[3], gs[3];
	int g[3];
	int ncell;
};

struct LloydParams {
	const float *f0, *f1, *f2;  // FM_F32 planes
	const uint32_t *rgba;       // FM_RGBA8 pixels (any packed 4 x u8 pixel: RGBA or HSVA)
	const float *lut3;          // FM_RGBA8: 3 x 256 fp32 feature tables per byte (null = identity)
	int mask_mode;              // FM_RGBA8: 0 = b0+b1+b2 > min_rgb_sum, 1 = b2 > min_rgb_sum
	long long n;
	int min_rgb_sum;
	double x2max;  // caller's bound on |x|^2 over all pixels (integer-key scheme)
	const double *centers;  // K x 3 fp64
	int K;
	uint32_t keymask;  // ~(KP-1); passed at run time so (key & mask) | idx stays one LOP3
	uint8_t *labels;
	double *sums, *counts, *inertia;
	double *partials;              // (unused by the kernel; scratch of other kernels)
	unsigned long long *pwords;    // per-CTA partials as self-validating words, see cs_ctx::d_partial_words
	unsigned long long launch_epoch;
	unsigned int *counter;
	double *centers_out, *stats;  // fused finalize (nullable)
	// multi-GPU exchange over peer memory (world == 1: unused)
	cs_mailbox *mb[kMgMaxRanks];
	int world, rank;
	unsigned long long epoch;
	// batched launch (gridDim.y = images, all of n pixels): element strides between consecutive images
	long long img_stride_px;     // pixels (features / packed pixels)
	long long img_stride_label;  // bytes of the label map
	// device-side loop control (cs_lloyd_run_*, nullable): ctl[0] = halt (0 run, 1 converged, 2 an empty
	// cluster needs the host), ctl[1] = iterations completed, ctl[2] = tol.  A launch that finds halt != 0
	// returns at once, so the host can queue a batch of iterations without synchronising in between.
	double *ctl;
	GridGeom grid;             // GRID kernels: cell grid over the caller's feature box
	const uint32_t *grid_tab;  // GRID kernels: kGridWords u32 written by grid_build_kernel for these centres
	uint32_t *grid_marks;      // GRID kernels: one bit per cell, set when a pixel of the cell took the all-K walk (cs_common.cuh)
	unsigned long long *phase_ts;  // CS_PHASE_TIMING builds: globaltimer stamps of block 0 (development)
};

template <int KP, int FM, class V, bool GRID = false> struct Smem {
	static constexpr int kPlanes = FM == FM_F32 ? 3 : 1;
	static constexpr int kStageBytes = kPlanes * V::TILE * 4;
	static constexpr int kAccBytes = V::NW * KP * KCfg<KP>::kCopies * 16;
	// GRID: kTabCopies copies of every 16-byte centre entry, copy j in bank group j (see grid_key)
	static constexpr int kTabCopies = GRID ? KCfg<KP>::kGridTabCopies : 1;
	static constexpr int kTabBytes = KP * 16 * kTabCopies;  // (non-GRID: KP/2 pairs x 2 float4)
	static constexpr int kC64Bytes = KP * 3 * 8;  // fp64 centres for the exact re-evaluation
	static constexpr int kRedBytes = (KP * 4 + 32) * 8;
	static constexpr int kLutBytes = FM == FM_RGBA8 ? 3 * 256 * 4 : 0;
	static constexpr int kGridBytes = GRID ? (KCfg<KP>::kGridCapUsed + 2 * KCfg<KP>::kGridPoolUsed) * 4 : 0;  // candidate table + overflow pool
	static constexpr int kFixed = kAccBytes + kTabBytes + kC64Bytes + kRedBytes + kLutBytes + kGridBytes + 128;
	static constexpr int kFit = (kSmemBudget - kFixed) / kStageBytes;
	static constexpr int kStages = kFit > 4 ? 4 : kFit;
	static_assert(kStages >= 2, "shared-memory ring needs at least 2 stages");
	static constexpr int kRingBytes = kStages * kStageBytes;
	static constexpr int kOffAcc = kRingBytes;
	static constexpr int kOffTab = kOffAcc + kAccBytes;
	static constexpr int kOffC64 = kOffTab + kTabBytes;
	static constexpr int kOffRed = kOffC64 + kC64Bytes;
	static constexpr int kOffLut = kOffRed + kRedBytes;
	static constexpr int kOffGrid = kOffLut + kLutBytes;
	static constexpr int kOffBar = kOffGrid + kGridBytes;
	static constexpr int kTotal = kOffBar + 2 * kStages * 8 + 16;  // (+ one more barrier for the grid table copy)
};

// ---- M-step tail (sklearn/cluster/_k_means_common.pyx:274-311, _kmeans.py:731-738) -----
// numpy's pairwise summation of a contiguous fp64 vector (n <= 256), reproduced so that
// (center_shift**2).sum() compares against tol exactly as in _kmeans_single_lloyd.
__device__ double np_pairwise_sum(const double *a, int n) {
	if (n < 8) {
		double r = 0.0;
		for (int i = 0; i < n; ++i) r += a[i];
		return r;
	}
	if (n <= 128) {
		double r[8];
		for (int j = 0; j < 8; ++j) r[j] = a[j];
		int i = 8;
		for (; i < n - (n % 8); i += 8)
			for (int j = 0; j < 8; ++j) r[j] += a[i + j];
		double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
		for (; i < n; ++i) res += a[i];
		return res;
	}
	int n2 = n / 2;
	n2 -= n2 % 8;
	return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// Runs on one CTA of >= max(K, 32) threads (K <= CS_MAX_K = 256 <= blockDim.x): thread k owns centre k.
// `shift2` is K doubles of shared scratch.  sums / counts / c_old may live in shared or global memory
// (already visible to this CTA); the fused caller hands in shared copies so that nothing on this
// serial tail waits for an L2 round trip.  Thread 0 returns the tol-test inputs in o_shift2 / o_nempty.
__device__ void finalize_block(const double *sums, const double *counts, const double *c_old,
                               int K, double *c_new, double *stats, double *shift2,
                               double &o_shift2, int &o_nempty) {
	const int t = threadIdx.x;
	__shared__ int s_argmax, s_nempty;
	__shared__ double s_total;
	if (t < 32) {
		// np.argmax (FIRST maximum), number of empty clusters, total weight: lane-strided scan, then a
		// butterfly that prefers the larger weight and, on equal weights, the lower index
		double best = -1.0, tot = 0.0;
		int am = 0x7fffffff, ne = 0;
		for (int k = t; k < K; k += 32) {
			const double w = counts[k];
			tot += w;  // integer-valued weights: exact in any order
			if (w > best) { best = w; am = k; }
			if (w == 0.0) ++ne;
		}
		for (int o = 16; o > 0; o >>= 1) {
			const double ob = __shfl_xor_sync(0xffffffffu, best, o);
			const int oa = __shfl_xor_sync(0xffffffffu, am, o);
			tot += __shfl_xor_sync(0xffffffffu, tot, o);
			ne += __shfl_xor_sync(0xffffffffu, ne, o);
			if (ob > best || (ob == best && oa < am)) { best = ob; am = oa; }
		}
		if (t == 0) { s_argmax = am; s_nempty = ne; s_total = tot; }
	}
	const int k = t;
	const bool mine = k < K;
	const double w = mine ? counts[k] : 1.0;
	double c0 = 0.0, c1 = 0.0, c2 = 0.0;
	if (mine && w > 0.0) {
		const double alpha = 1.0 / w;  // _average_centers: alpha = 1/w, centre *= alpha
		c0 = sums[3 * k + 0] * alpha; c1 = sums[3 * k + 1] * alpha; c2 = sums[3 * k + 2] * alpha;
		c_new[3 * k + 0] = c0; c_new[3 * k + 1] = c1; c_new[3 * k + 2] = c2;
	}
	__syncthreads();
	if (s_nempty > 0) {  // CTA-uniform
		if (t == 0) {
			// _average_centers walks j in order and copies centers[argmax] *as it is at that
			// moment*: still the raw sum for j < argmax, the averaged centre for j > argmax.
			const int am = s_argmax;
			for (int e = 0; e < K; ++e) {
				if (counts[e] > 0.0) continue;
				for (int j = 0; j < 3; ++j)
					c_new[3 * e + j] = (e < am) ? sums[3 * am + j] : c_new[3 * am + j];
			}
		}
		__syncthreads();
		if (mine && !(w > 0.0)) { c0 = c_new[3 * k + 0]; c1 = c_new[3 * k + 1]; c2 = c_new[3 * k + 2]; }
	}
	if (mine) {
		// _euclidean_dense_dense, n_features = 3: sequential remainder loop
		double s = 0.0, d;
		// (separate multiply and add, as the C loop compiles on x86-64: no fused multiply-add)
		d = c0 - c_old[3 * k + 0]; s = __dadd_rn(s, __dmul_rn(d, d));
		d = c1 - c_old[3 * k + 1]; s = __dadd_rn(s, __dmul_rn(d, d));
		d = c2 - c_old[3 * k + 2]; s = __dadd_rn(s, __dmul_rn(d, d));
		const double sh = sqrt(s);  // center_shift[k]
		shift2[k] = sh * sh;        // (center_shift ** 2)
	}
	__syncthreads();
	if (t == 0) {
		o_shift2 = np_pairwise_sum(shift2, K);
		o_nempty = s_nempty;
		stats[0] = o_shift2;
		stats[1] = (double)s_nempty;
		stats[2] = (double)s_argmax;
		stats[3] = s_total;
	}
}

// fp64 re-evaluation of one pixel: first minimum of the direct squared distance.
__device__ __noinline__ int exact_label(float x, float y, float z, const double *c64, int K) {
	double best = 1e300;
	int bi = 0;
	for (int k = 0; k < K; ++k) {
		double dx = (double)x - c64[3 * k], dy = (double)y - c64[3 * k + 1],
		       dz = (double)z - c64[3 * k + 2];
		double d = dx * dx + dy * dy + dz * dz;
		if (d < best) { best = d; bi = k; }
	}
	return bi;
}

// Self-validating words: a double travels as two scalar 8-byte words (tag << 32) | 32 data bits, each written
// with ONE single-copy-atomic store, so a reader that finds the expected tag has the data — no flag, no fence.
// Used between the CTAs of a launch (tag = launch epoch, gpu scope) and between the GPUs of a sharded run
// (tag = exchange epoch, system scope over NVLink P2P).
__device__ __forceinline__ void word_store_gpu(unsigned long long *p, uint32_t data, uint32_t tag) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)tag << 32) | data) : "memory");
}
__device__ __forceinline__ unsigned long long word_load_gpu(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void word_store_sys(unsigned long long *p, uint32_t data, uint32_t tag) {
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)tag << 32) | data) : "memory");
}
__device__ __forceinline__ unsigned long long word_load_sys(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void put_double_gpu(unsigned long long *w, double v, uint32_t tag) {
	const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
	word_store_gpu(w, (uint32_t)bits, tag);
	word_store_gpu(w + 1, (uint32_t)(bits >> 32), tag);
}
// waits (bounded: 2 s, then NaN and *timed_out = 1) until both words of a double carry `tag`
__device__ __forceinline__ double get_double_gpu(const unsigned long long *w, uint32_t tag, int *timed_out) {
	unsigned long long lo = word_load_gpu(w), hi = word_load_gpu(w + 1);
	if ((uint32_t)(lo >> 32) != tag || (uint32_t)(hi >> 32) != tag) {
		unsigned long long t0;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		do {
			unsigned long long t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if (t1 - t0 > 2000000000ull) { *timed_out = 1; return __longlong_as_double(0x7ff8000000000000ll); }
			lo = word_load_gpu(w); hi = word_load_gpu(w + 1);
		} while ((uint32_t)(lo >> 32) != tag || (uint32_t)(hi >> 32) != tag);
	}
	return __longlong_as_double((long long)((hi << 32) | (lo & 0xFFFFFFFFull)));
}
__device__ __forceinline__ unsigned long long mg_globaltimer() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// Shared (CTA-uniform) constants of the key scheme, written once in the prologue.
struct KeyConst {
	float S;       // power-of-two scale of the integer-key scheme
	float x2max;   // bound on |x|^2 handed in by the caller
	float tau_v;   // near-tie threshold in scaled-key units (integer-key scheme)
	float cnmax;   // max |c|^2 (float-key scheme)
	uint32_t sh;   // 1 << bits, kept in a register so (bits(v) * sh + addend) is ONE IMAD
	float pad;     // key value of a padding table entry (top of the window)
};

// Distances + argmin + (exact re-evaluation) + accumulation for the P pixels of one consumer
// thread.  FULL = every pixel is real and unmasked (warp-uniform fast path).
//
// Key scheme for KP <= 64 ("integer keys"): with v = (|c|^2 - 2 x.c + x2max) S + 1
// >= 1, the IEEE bit pattern of v is monotone in v, so  key = (bits(v) - bits(1.0f)) << b | k
// orders by distance first and centre index second with NO mantissa bits lost; it is one IMAD
// (bits(v) * 2^b + const_k, mod 2^32) and the argmin is a chain of 2-input unsigned mins —
// on sm_100 FFMA, IADD and 2-input min/max issue every cycle while LOP3, IMAD, 3-input min and
// the packed f32x2 ops take two, and nothing overlaps (tools/ubench/pipes.cu, profiles/).
// |x|^2 is not added per centre: it does not change the order.
// Key scheme for KP >= 128 ("float keys"): d = |x|^2 + |c|^2 - 2 x.c with the index in the low
// mantissa bits (the 2^(32-b) window of the integer scheme is too narrow for b >= 7).
template <int KP, int FM, bool TIE, bool INERTIA, class V, bool FULL, int P>
__device__ __forceinline__ void assign_update(
    const float (&x)[P], const float (&y)[P], const float (&z)[P], const bool (&use)[P],
    int (&lab)[P], const uint32_t tab_s, const float2 (&treg)[KP <= 16 ? KP * 2 : 1],
    const double *c64, int K, uint32_t keymask, const KeyConst &kc, float4 *wacc, int lane,
    float &inert) {
	constexpr int kCopies = KCfg<KP>::kCopies, kPhases = KCfg<KP>::kPhases;
	constexpr int kBits = KCfg<KP>::kBits;
	constexpr bool INTKEY = KP <= 64;
	float xx[P];
	uint32_t ibest[P], isecond[P];
	float fbest[P], fsecond[P];
#pragma unroll
	for (int q = 0; q < P; ++q) {
		if (!INTKEY || INERTIA) xx[q] = fmaf(x[q], x[q], fmaf(y[q], y[q], z[q] * z[q]));
		ibest[q] = 0xFFFFFFFFu; isecond[q] = 0xFFFFFFFFu;
		fbest[q] = 3.0e38f; fsecond[q] = 3.0e38f;
	}
	constexpr uint32_t kBias = 0x3F800000u << kBits;  // bits(1.0f) << b, mod 2^32
	constexpr int PPASS = (KP <= 16 && P % V::PP == 0) ? V::PP : P;
	// labels + slot update once, after the last pass (doing it after every pass measured slower: sweep5)
	constexpr int UPASS = P;
	char *wslot = reinterpret_cast<char *>(wacc + (lane % kCopies));
#pragma unroll
	for (int q0 = 0; q0 < P; q0 += PPASS) {
#pragma unroll(KP <= 16 ? KP / 2 : 4)
	for (int pr = 0; pr < KP / 2; ++pr) {
		float2 mx, my, mz, cn;
		if (V::TABREG && KP <= 16) {
			mx = treg[4 * pr]; my = treg[4 * pr + 1]; mz = treg[4 * pr + 2]; cn = treg[4 * pr + 3];
		} else {
			// (explicit shared-space address: the generic pointer made ptxas rebuild the shared window base
			// from SR_CgaCtaId in every tile iteration)
			const float4 t0 = lds128(tab_s + pr * 32), t1 = lds128(tab_s + pr * 32 + 16);
			mx = make_float2(t0.x, t0.y); my = make_float2(t0.z, t0.w);
			mz = make_float2(t1.x, t1.y); cn = make_float2(t1.z, t1.w);
		}
		const uint32_t add0 = (uint32_t)(2 * pr) - kBias, add1 = (uint32_t)(2 * pr + 1) - kBias;
#pragma unroll
		for (int q = q0; q < q0 + PPASS; ++q) {
			if (INTKEY) {
				float2 d = __ffma2_rn(make_float2(z[q], z[q]), mz, cn);
				d = __ffma2_rn(make_float2(y[q], y[q]), my, d);
				d = __ffma2_rn(make_float2(x[q], x[q]), mx, d);
				// shift by a constant (LEA); a multiply by a run-time 2^b (IMAD) measured 15 % slower in the
				// core-loop lab (tools/ubench/core.cu)
				const uint32_t k0 = (__float_as_uint(d.x) << kBits) + add0;
				const uint32_t k1 = (__float_as_uint(d.y) << kBits) + add1;
				if (TIE) {
					// second smallest of {best, second, k0, k1} = min(second, max(best, lo), hi)
					const uint32_t lo = min(k0, k1), hi = max(k0, k1);
					isecond[q] = min(min(isecond[q], max(ibest[q], lo)), hi);
					ibest[q] = min(ibest[q], lo);
				} else if (V::MINM == 1) {
					ibest[q] = min(ibest[q], k0);
					asm volatile("" : "+r"(ibest[q]));
					ibest[q] = min(ibest[q], k1);
				} else {
					ibest[q] = min(min(ibest[q], k0), k1);
				}
			} else {
				float2 d = __fadd2_rn(cn, make_float2(xx[q], xx[q]));
				d = __ffma2_rn(make_float2(z[q], z[q]), mz, d);
				d = __ffma2_rn(make_float2(y[q], y[q]), my, d);
				d = __ffma2_rn(make_float2(x[q], x[q]), mx, d);
				const float k0 = __uint_as_float((__float_as_uint(d.x) & keymask) | (uint32_t)(2 * pr));
				const float k1 = __uint_as_float((__float_as_uint(d.y) & keymask) | (uint32_t)(2 * pr + 1));
				if (TIE) {
					const float lo = fminf(k0, k1), hi = fmaxf(k0, k1);
					fsecond[q] = fminf(fminf(fsecond[q], fmaxf(fbest[q], lo)), hi);
					fbest[q] = fminf(fbest[q], lo);
				} else {
					fbest[q] = fminf(fminf(fbest[q], k0), k1);
				}
			}
		}
	}
	// ---- labels + update ----
	if ((q0 + PPASS) % UPASS != 0) continue;
	const int u0 = q0 + PPASS - UPASS;
#pragma unroll
	for (int q = u0; q < u0 + UPASS; ++q) {
		float dbest;  // squared distance to the fp32 winner (for the inertia)
		if (INTKEY) {
			lab[q] = (int)(ibest[q] & (uint32_t)(KP - 1));
			const float vb = __uint_as_float((ibest[q] >> kBits) + 0x3F800000u);
			if (TIE) {
				const float vs = __uint_as_float((isecond[q] >> kBits) + 0x3F800000u);
				if ((FULL || use[q]) && (vs - vb) <= kc.tau_v) lab[q] = exact_label(x[q], y[q], z[q], c64, K);
			}
			if (INERTIA) dbest = (vb - 1.0f) / kc.S - kc.x2max + xx[q];
		} else {
			lab[q] = (int)(__float_as_uint(fbest[q]) & (uint32_t)(KP - 1));
			if (TIE) {
				// |key - d| <= A (|x|^2 + 2 max|c|^2) + B d,  A = 5*2^-24, B = 2^-(23-bits)
				const float tauA = 2.2f * 5.0f * 5.9604645e-8f;
				const float tauB = 2.2f / (float)(1 << (23 - kBits));
				const float tau = fmaf(tauB, fmaxf(fsecond[q], 0.f), tauA * (xx[q] + 2.f * kc.cnmax));
				if ((FULL || use[q]) && (fsecond[q] - fbest[q]) <= tau) lab[q] = exact_label(x[q], y[q], z[q], c64, K);
			}
			if (INERTIA) dbest = __uint_as_float(__float_as_uint(fbest[q]) & keymask);
		}
		if (INERTIA && (FULL || use[q])) {
			if (TIE) {  // distance to the label actually taken
				const float dx = x[q] - (float)c64[3 * lab[q]], dy = y[q] - (float)c64[3 * lab[q] + 1],
				            dz = z[q] - (float)c64[3 * lab[q] + 2];
				dbest = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
			}
			inert += fmaxf(dbest, 0.f);
		}
	}
	// update: lane-private slots; kPhases groups of kCopies lanes take turns.
	// The label is made opaque to the optimiser first: otherwise it folds `key & (KP-1)` into the slot
	// address as shift + and + or + add (four half-rate instructions per pixel); this way the address
	// is ONE shift-add on the label that the label store needs anyway.
#pragma unroll
	for (int q = u0; q < u0 + UPASS; ++q) asm volatile("" : "+r"(lab[q]));
#pragma unroll
	for (int ph = 0; ph < kPhases; ++ph) {
		if (kPhases == 1 || (lane / kCopies) == ph) {
#pragma unroll
			for (int q = u0; q < u0 + UPASS; ++q) {
				if (FULL || use[q]) {
					float4 *slot = reinterpret_cast<float4 *>(wslot + (uint32_t)lab[q] * (uint32_t)(kCopies * 16));
					float4 v = *slot;
					v.x += x[q]; v.y += y[q]; v.z += z[q]; v.w += 1.f;
					*slot = v;
				}
			}
		}
		if (kPhases > 1) __syncwarp();
	}
	}  // passes
}

// ================= grid-filtered exact assignment (GRID kernels; 4 <= K <= 64, planar fp32) =================
// The nearest centre of a pixel can only be one of the centres that are NOT dominated over the pixel's
// cell of a regular grid over feature space (centre k is dominated when some centre w is closer than k at
// every point of the cell — the filtering test of Kanungo et al.'s kd-tree k-means, on a uniform grid).
// grid_build_kernel lists, once per Lloyd iteration, the candidates of every cell; the Lloyd kernel copies
// the table (<= 46 KB) into shared memory and evaluates FOUR distances per pixel instead of K (four more for
// the few pixels of cells with five to eight candidates), then the usual test: when the two best of them are
// closer than the rounding bound of the keys the pixel is re-evaluated in fp64 over the cell's candidates —
// so the label is the fp64 first minimum (the oracle's), as in the full walk with CS_LLOYD_EXACT_TIES.  Every
// step is sound for ANY input: border cells extend to infinity, cell boxes are inflated by more than the
// rounding of the index arithmetic, padding candidates are real distinct centres, a centre is dropped from
// a cell only when another one is strictly closer everywhere in it, cells with more than eight candidates
// (or without a pool entry) walk over all K.  The per-pixel cost hardly depends on K — but every lane gathers
// four DIFFERENT 16-byte centre entries, which makes the kernel shared-memory-bound (~31 wavefronts per 32
// pixels, 86 % of the peak rate): slower than the full walk up to K = 8, faster from K = 9 on shards of 10 MP
// and more (profiles/r2_grid_assignment.md).
//
// Table entry (u32), 16 < K <= 64: four label bytes, ascending, so that the slot index in the low key bits
// breaks equal distances towards the lowest label.  Overflow entries carry byte0 > byte1 = 0: byte0 = 1 ->
// pool index in bytes 2 (low log2 KP bits) and 3 (the rest); byte0 = 2 -> all K centres.  K <= 16: eight
// nibbles, see assign_grid_nib.
//
// Keys are FIXED-POINT here: t = (|c|^2 - 2 x.c + x2max) s + 1.5 * 2^23 is formed by the three FMAs
// themselves (every partial sum stays inside [2^23, 2^24), where the fp32 spacing is 1), so bits(t) is
// linear in the squared distance and the near-tie test is one integer subtraction.
constexpr float kGridMagic = 12582912.0f;      // 1.5 * 2^23
constexpr uint32_t kGridKeyBase = 0x4B000000u;  // bits(2^23)
constexpr int kGridTauD = 8;  // 3 FMA roundings (<= 1/2 unit each) + 4 rounded table entries, two keys, with slack

struct GridConst {
	float sx, ox, gx, sy, oy, gy, sz, oz, gz;
	uint32_t stride_y, stride_z;  // g[0], g[0] * g[1]
	uint32_t base_c;              // shared address of the table minus 4 * bits(magic) * (1 + stride_y + stride_z)
	uint32_t pool_s, ctab_s;
	uint32_t logkp;
	uint32_t *marks;  // see LloydParams::grid_marks
};

__device__ __forceinline__ float ffma_sat(float a, float b, float c) {
	float d;
	asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ float ffma_rm(float a, float b, float c) {
	float d;
	asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
	return d;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// fixed-point key of centre `label` for pixel (x,y,z), slot index (0..7) in the low 3 bits.  Entry of centre l at
// ctab_s + (l << SH).  SH = 4: one 16-byte entry per centre; the 8 lanes of a quarter-warp (one LDS.128 phase)
// gather different entries and collide whenever two of them share a 16-byte bank group.  SH = 7: eight copies
// per centre, copy j in bank group j, and ctab_s already holds the lane's (lane & 7) * 16 — every phase is
// conflict-free whatever the labels are (4 wavefronts per LDS.128 instead of up to 7 measured at K = 16).
template <int SH>
__device__ __forceinline__ uint32_t grid_key3(float x, float y, float z, uint32_t ctab_s, uint32_t label, uint32_t slot) {
	const float4 t = lds128(ctab_s + (label << SH));
	const float v = fmaf(x, t.x, fmaf(y, t.y, fmaf(z, t.z, t.w)));
	return ((__float_as_uint(v) - kGridKeyBase) << 3) + slot;
}

// rare path: all K centres in fp32 (same fixed-point keys, index in the low 8 bits), fp64 only on a near tie
template <int SH>
__device__ __noinline__ int grid_walk_all_label(float x, float y, float z, uint32_t ctab_s, const double *c64, int K) {
	uint32_t best = 0xFFFFFFFFu, sec = 0xFFFFFFFFu;
	for (int k = 0; k < K; ++k) {
		const float4 t = lds128(ctab_s + ((uint32_t)k << SH));
		const float v = fmaf(x, t.x, fmaf(y, t.y, fmaf(z, t.z, t.w)));
		const uint32_t key = ((__float_as_uint(v) - kGridKeyBase) << 8) + (uint32_t)k;
		sec = min(sec, max(best, key));
		best = min(best, key);
	}
	if ((sec >> 8) - (best >> 8) > (uint32_t)kGridTauD) return (int)(best & 0xFFu);
	return exact_label(x, y, z, c64, K);
}

// Rare paths of assign_grid.  Both are INLINE on purpose: a __noinline__ callee saves its ~48 callee-saved registers
// to local memory on entry and reloads them on return, and with 227 KB of the SM's L1 configured as shared memory
// that stack traffic goes to L2 — measured on the B200 (64 MP, uniform-random image): the lane-serial callee cost
// ~1000 cycles per visit, and with 0.2 % / 1 % / 3 % of the pixels in pool cells at K = 16 / 32 / 64, i.e. a visit
// in 23 % / 72 % / 98 % of a warp's passes, the handling of those pixels took 0.034 / 0.126 / 0.30 ms of the
// 0.302 / 0.430 / 0.766 ms iteration (profiles/r2_grid_assignment.md, "rare-path cost").

// fp64 first minimum over the candidates of one table word (ascending labels, so the strict < keeps the lowest
// label among equal distances).  Sound because the true nearest centre — and every centre at exactly its
// distance — is among the cell's candidates (a centre is dropped only when another one is strictly closer
// everywhere in the inflated cell, grid_build_kernel), and the padding entries are real centres.
__device__ __forceinline__ void exact_word(float x, float y, float z, uint32_t w, const double *c64, double &best, int &bi) {
#pragma unroll
	for (int s = 0; s < 4; ++s) {
		const int l = (int)((w >> (8 * s)) & 0xFFu);
		const double dx = (double)x - c64[3 * l], dy = (double)y - c64[3 * l + 1], dz = (double)z - c64[3 * l + 2];
		const double d = dx * dx + dy * dy + dz * dz;
		if (d < best) { best = d; bi = l; }
	}
}

// a pixel that takes the all-K walk marks its cell, so that the next table build gives the cell a pool entry first
__device__ __forceinline__ void grid_mark_cell(float x, float y, float z, const GridConst &gc) {
	const float ux = ffma_sat(x, gc.sx, gc.ox), uy = ffma_sat(y, gc.sy, gc.oy), uz = ffma_sat(z, gc.sz, gc.oz);
	const uint32_t ix = __float_as_uint(ffma_rm(ux, gc.gx, kGridMagic)) - __float_as_uint(kGridMagic);
	const uint32_t iy = __float_as_uint(ffma_rm(uy, gc.gy, kGridMagic)) - __float_as_uint(kGridMagic);
	const uint32_t iz = __float_as_uint(ffma_rm(uz, gc.gz, kGridMagic)) - __float_as_uint(kGridMagic);
	const uint32_t cell = iz * gc.stride_z + (iy * gc.stride_y + ix);
	if (cell < (uint32_t)kGridCap) atomicOr(gc.marks + (cell >> 5), 1u << (cell & 31u));
}

// labels for the P pixels of one consumer thread through the cell table (no accumulation); 16 < K <= 64.
// A pixel of a pool cell (five to eight candidates) takes the FIRST word of its pool entry in place of the cell entry
// before the pass — one more load for those lanes only — so the pass evaluates candidates 0..3 for it like for
// everybody else; the rare path then evaluates only candidates 4..7 and merges.  Keys carry the slot (0..7) in
// their low three bits: equal distances go to the lowest slot = the lowest label.
template <bool FULL, int P, int SH, bool MARK>
__device__ __forceinline__ void assign_grid(const float (&x)[P], const float (&y)[P], const float (&z)[P],
                                            const bool (&use)[P], int (&lab)[P], const GridConst &gc,
                                            const double *c64, int K) {
	uint32_t eo[P], e[P], best[P], sec[P];
#pragma unroll
	for (int q = 0; q < P; ++q) {
		const float ux = ffma_sat(x[q], gc.sx, gc.ox), uy = ffma_sat(y[q], gc.sy, gc.oy), uz = ffma_sat(z[q], gc.sz, gc.oz);
		const uint32_t ix = __float_as_uint(ffma_rm(ux, gc.gx, kGridMagic));  // bits(magic) + floor(u * g)
		const uint32_t iy = __float_as_uint(ffma_rm(uy, gc.gy, kGridMagic));
		const uint32_t iz = __float_as_uint(ffma_rm(uz, gc.gz, kGridMagic));
		const uint32_t cell = iz * gc.stride_z + (iy * gc.stride_y + ix);
		eo[q] = lds32((cell << 2) + gc.base_c);
	}
#pragma unroll
	for (int q = 0; q < P; ++q) {
		e[q] = eo[q];
		if ((eo[q] & 0xFFFFu) == 1u) {  // pool cell: byte0 = 1, byte1 = 0, pool index in bytes 2 and 3
			const uint32_t pidx = ((eo[q] >> 16) & 0xFFu) | ((eo[q] >> 24) << gc.logkp);
			e[q] = lds32(gc.pool_s + pidx * 8u);
		}
	}
	bool rare[P];
	bool any_rare = false;
#pragma unroll
	for (int q = 0; q < P; ++q) {
		const uint32_t l0 = __byte_perm(e[q], 0u, 0x4440), l1 = __byte_perm(e[q], 0u, 0x4441);
		const uint32_t l2 = __byte_perm(e[q], 0u, 0x4442), l3 = __byte_perm(e[q], 0u, 0x4443);
		const uint32_t k0 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l0, 0u), k1 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l1, 1u);
		const uint32_t k2 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l2, 2u), k3 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l3, 3u);
		const uint32_t a = min(k0, k1), A = max(k0, k1), b = min(k2, k3), B = max(k2, k3);
		best[q] = min(a, b);
		sec[q] = min(max(a, b), min(A, B));
		// label of the winning slot: byte (best & 3) of the entry
		lab[q] = (int)__byte_perm(e[q], 0u, (best[q] & 3u) | 0x4440u);
		// rare: an overflow cell (byte1 = 0), or the two best keys closer than their rounding bound
		rare[q] = (FULL || use[q]) && ((eo[q] & 0xFF00u) == 0u || (sec[q] - best[q]) <= (uint32_t)(8 * kGridTauD + 7));
		any_rare = any_rare || rare[q];
	}
#ifdef CS_GRID_NORARE
	if (false) {
#else
	if (any_rare) {
#endif
#pragma unroll
		for (int q = 0; q < P; ++q) {
			if (rare[q]) {
				double bd = 1e300;
				int bi = 0;
				if ((eo[q] & 0xFF00u) == 0u) {
					if ((eo[q] & 0xFFu) != 1u) {  // more than eight candidates (or the pool is full): all K centres
						lab[q] = grid_walk_all_label<SH>(x[q], y[q], z[q], gc.ctab_s, c64, K);
						if constexpr (MARK) grid_mark_cell(x[q], y[q], z[q], gc);
						continue;
					}
					const uint32_t pidx = ((eo[q] >> 16) & 0xFFu) | ((eo[q] >> 24) << gc.logkp);
					const uint32_t wB = lds32(gc.pool_s + pidx * 8u + 4u);
					const uint32_t l4 = __byte_perm(wB, 0u, 0x4440), l5 = __byte_perm(wB, 0u, 0x4441);
					const uint32_t l6 = __byte_perm(wB, 0u, 0x4442), l7 = __byte_perm(wB, 0u, 0x4443);
					const uint32_t k4 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l4, 4u), k5 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l5, 5u);
					const uint32_t k6 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l6, 6u), k7 = grid_key3<SH>(x[q], y[q], z[q], gc.ctab_s, l7, 7u);
					const uint32_t a = min(k4, k5), A = max(k4, k5), b = min(k6, k7), B = max(k6, k7);
					const uint32_t bestB = min(a, b), secB = min(max(a, b), min(A, B));
					const uint32_t s8 = min(max(best[q], bestB), min(sec[q], secB));
					const uint32_t b8 = min(best[q], bestB);
					lab[q] = (int)(__byte_perm(e[q], wB, b8 & 7u) & 0xFFu);
					if ((s8 >> 3) - (b8 >> 3) > (uint32_t)kGridTauD) continue;
					exact_word(x[q], y[q], z[q], e[q], c64, bd, bi);
					exact_word(x[q], y[q], z[q], wB, c64, bd, bi);
				} else {  // the two best of the four candidates are closer than the rounding bound of their keys
					exact_word(x[q], y[q], z[q], e[q], c64, bd, bi);
				}
				lab[q] = bi;
			}
		}
	}
}

// ---- K <= 16: EIGHT candidates per cell, one nibble each, and no pool ----
// Entry (u32): nibbles n0..n7.  Up to four candidates: n0 < n1 < n2 < n3 (padded with real, distinct centres),
// n4..n7 = 0 — the high half is zero.  Five to eight candidates: eight distinct labels, ascending (padded likewise;
// needs K >= 8), so the high half is not zero: the pass evaluates n0..n3 as always, and only the lanes of such a
// pixel evaluate n4..n7 as well and merge (four gathers, no pool fetch, nothing evaluated twice).  0xFFFF0000 | x:
// more than eight candidates -> all K centres.  Keys carry 4 * slot in their low five bits, which is also the
// shift that extracts the slot's nibble; equal distances go to the lowest slot = the lowest label.
__device__ __forceinline__ uint32_t shr_wrap(uint32_t v, uint32_t sh) {
	uint32_t r;
	asm("shf.r.wrap.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(v), "r"(sh));
	return r;
}
template <int SH>
__device__ __forceinline__ uint32_t grid_key5(float x, float y, float z, uint32_t addr, uint32_t slot4) {
	const float4 t = lds128(addr);
	const float v = fmaf(x, t.x, fmaf(y, t.y, fmaf(z, t.z, t.w)));
	return ((__float_as_uint(v) - kGridKeyBase) << 5) + slot4;
}
// fp64 first minimum over `cnt` nibbles of w starting at nibble `first` (ascending labels: strict < keeps the lowest)
__device__ __forceinline__ void exact_nibbles(float x, float y, float z, uint32_t w, int first, int cnt, const double *c64,
                                              double &best, int &bi) {
	for (int s = first; s < first + cnt; ++s) {
		const int l = (int)((w >> (4 * s)) & 15u);
		const double dx = (double)x - c64[3 * l], dy = (double)y - c64[3 * l + 1], dz = (double)z - c64[3 * l + 2];
		const double d = dx * dx + dy * dy + dz * dz;
		if (d < best) { best = d; bi = l; }
	}
}
template <bool FULL, int P, int SH>
__device__ __forceinline__ void assign_grid_nib(const float (&x)[P], const float (&y)[P], const float (&z)[P],
                                                const bool (&use)[P], int (&lab)[P], const GridConst &gc,
                                                const double *c64, int K) {
	static_assert(SH == 7, "nibble entries: eight copies of the centre table");
	constexpr uint32_t kM = 15u << SH;
	uint32_t e[P], best[P], sec[P];
#pragma unroll
	for (int q = 0; q < P; ++q) {
		const float ux = ffma_sat(x[q], gc.sx, gc.ox), uy = ffma_sat(y[q], gc.sy, gc.oy), uz = ffma_sat(z[q], gc.sz, gc.oz);
		const uint32_t ix = __float_as_uint(ffma_rm(ux, gc.gx, kGridMagic));
		const uint32_t iy = __float_as_uint(ffma_rm(uy, gc.gy, kGridMagic));
		const uint32_t iz = __float_as_uint(ffma_rm(uz, gc.gz, kGridMagic));
		const uint32_t cell = iz * gc.stride_z + (iy * gc.stride_y + ix);
		e[q] = lds32((cell << 2) + gc.base_c);
	}
	bool rare[P];
	bool any_rare = false;
#pragma unroll
	for (int q = 0; q < P; ++q) {
		// byte offset of nibble s in the centre table: ((e >> 4s) & 15) << 7
		const uint32_t a0 = gc.ctab_s + ((e[q] << 7) & kM), a1 = gc.ctab_s + ((e[q] << 3) & kM);
		const uint32_t a2 = gc.ctab_s + ((e[q] >> 1) & kM), a3 = gc.ctab_s + ((e[q] >> 5) & kM);
		const uint32_t k0 = grid_key5<SH>(x[q], y[q], z[q], a0, 0u), k1 = grid_key5<SH>(x[q], y[q], z[q], a1, 4u);
		const uint32_t k2 = grid_key5<SH>(x[q], y[q], z[q], a2, 8u), k3 = grid_key5<SH>(x[q], y[q], z[q], a3, 12u);
		const uint32_t a = min(k0, k1), A = max(k0, k1), b = min(k2, k3), B = max(k2, k3);
		best[q] = min(a, b);
		sec[q] = min(max(a, b), min(A, B));
		lab[q] = (int)(shr_wrap(e[q], best[q]) & 15u);
		// rare: more than four candidates (high half set), or the two best keys closer than their rounding bound
		rare[q] = (FULL || use[q]) && (e[q] > 0xFFFFu || (sec[q] - best[q]) <= (uint32_t)(32 * kGridTauD + 31));
		any_rare = any_rare || rare[q];
	}
#ifdef CS_GRID_NORARE
	if (false) {
#else
	if (any_rare) {
#endif
#pragma unroll
		for (int q = 0; q < P; ++q) {
			if (rare[q]) {
				int ncand = 4;
				if (e[q] > 0xFFFFu) {
					if ((e[q] >> 16) == 0xFFFFu) {  // more than eight candidates: all K centres
						lab[q] = grid_walk_all_label<SH>(x[q], y[q], z[q], gc.ctab_s, c64, K);
						continue;
					}
					const uint32_t a4 = gc.ctab_s + ((e[q] >> 9) & kM), a5 = gc.ctab_s + ((e[q] >> 13) & kM);
					const uint32_t a6 = gc.ctab_s + ((e[q] >> 17) & kM), a7 = gc.ctab_s + ((e[q] >> 21) & kM);
					const uint32_t k4 = grid_key5<SH>(x[q], y[q], z[q], a4, 16u), k5 = grid_key5<SH>(x[q], y[q], z[q], a5, 20u);
					const uint32_t k6 = grid_key5<SH>(x[q], y[q], z[q], a6, 24u), k7 = grid_key5<SH>(x[q], y[q], z[q], a7, 28u);
					const uint32_t a = min(k4, k5), A = max(k4, k5), b = min(k6, k7), B = max(k6, k7);
					const uint32_t bestB = min(a, b), secB = min(max(a, b), min(A, B));
					const uint32_t s8 = min(max(best[q], bestB), min(sec[q], secB));
					const uint32_t b8 = min(best[q], bestB);
					lab[q] = (int)(shr_wrap(e[q], b8) & 15u);
					if ((s8 >> 5) - (b8 >> 5) > (uint32_t)kGridTauD) continue;
					ncand = 8;
				}
				double bd = 1e300;
				int bi = 0;
				exact_nibbles(x[q], y[q], z[q], e[q], 0, ncand, c64, bd, bi);
				lab[q] = bi;
			}
		}
	}
}

// lane-private slot update shared by the GRID kernels (same scheme as at the end of assign_update)
template <int KP, bool FULL, int P>
__device__ __forceinline__ void update_slots(const float (&x)[P], const float (&y)[P], const float (&z)[P],
                                             const bool (&use)[P], int (&lab)[P], float4 *wacc, int lane) {
	constexpr int kCopies = KCfg<KP>::kCopies, kPhases = KCfg<KP>::kPhases;
	char *wslot = reinterpret_cast<char *>(wacc + (lane % kCopies));
#pragma unroll
	for (int q = 0; q < P; ++q) asm volatile("" : "+r"(lab[q]));
	if constexpr (kPhases == 4) {
		// K = 64: eight slot copies, so four lanes (lane, lane ^ 8, lane ^ 16, lane ^ 24) share one.  Instead of four
		// phases with a quarter of the lanes each, every lane learns with three shuffles how many of its LOWER
		// partners hold the same label for pixel q (its rank); round r serves the lanes of rank r.  With 64 labels
		// a pass of four pixels needs two rounds as a rule (a third in 2 % of the passes) instead of four phases —
		// and a phase, like a round, is four dependent read-modify-writes (the compiler cannot overlap accesses to
		// slots that may be the same).  The __syncwarp after every pixel keeps the order between a lane's store
		// for pixel q and a partner's load for pixel q + 1 of the same slot.
		int rank[P], rmax = 0;
#pragma unroll
		for (int q = 0; q < P; ++q) {
			const int l = (FULL || use[q]) ? lab[q] : -1 - lane;  // a pixel that is not accumulated conflicts with nobody
			const int a = __shfl_xor_sync(0xffffffffu, l, 8), b = __shfl_xor_sync(0xffffffffu, l, 16), c = __shfl_xor_sync(0xffffffffu, l, 24);
			rank[q] = (int)(a == l && (lane ^ 8) < lane) + (int)(b == l && (lane ^ 16) < lane) + (int)(c == l && (lane ^ 24) < lane);
			rmax = max(rmax, rank[q]);
		}
		rmax = (int)__reduce_max_sync(0xffffffffu, (unsigned)rmax);
		for (int r = 0; r <= rmax; ++r) {
#pragma unroll
			for (int q = 0; q < P; ++q) {
				if ((FULL || use[q]) && rank[q] == r) {
					float4 *slot = reinterpret_cast<float4 *>(wslot + (uint32_t)lab[q] * (uint32_t)(kCopies * 16));
					float4 v = *slot;
					v.x += x[q]; v.y += y[q]; v.z += z[q]; v.w += 1.f;
					*slot = v;
				}
				__syncwarp();
			}
		}
	} else {
#pragma unroll
		for (int ph = 0; ph < kPhases; ++ph) {
			if (kPhases == 1 || (lane / kCopies) == ph) {
#pragma unroll
				for (int q = 0; q < P; ++q) {
					if (FULL || use[q]) {
						float4 *slot = reinterpret_cast<float4 *>(wslot + (uint32_t)lab[q] * (uint32_t)(kCopies * 16));
						float4 v = *slot;
						v.x += x[q]; v.y += y[q]; v.z += z[q]; v.w += 1.f;
						*slot = v;
					}
				}
			}
			if (kPhases > 1) __syncwarp();
		}
	}
}

// Candidate table for the centres of the NEXT Lloyd launch on the stream.  One warp per cell, lanes stride over
// the centres: k is a candidate unless some centre w is closer at every point of the (inflated) cell box, i.e.
// max over the box of d_w - d_k = |c_w|^2 - |c_k|^2 + 2 x.(c_k - c_w) is negative.  First filter: the centre
// nearest to the middle of the box (the filtering algorithm's choice); then all pairs among the (<= 32)
// survivors.  fp64 throughout.
constexpr int kGridMaxK = 64;
#define CS_GB_BOUNDS __global__ void __launch_bounds__(256)  // (forced to 32 registers with one cell per warp: measured slower)
// HALF (K <= 16): a cell takes half a warp, so a warp builds two cells at a time — at K <= 16 only 16 lanes had work
// in the centre loops, and the kernel's length is the latency of the one or two cells each warp walks through.
template <bool HALF>
CS_GB_BOUNDS grid_build_kernel(const double *__restrict__ centers, int K, GridGeom g,
                                                         uint32_t *__restrict__ out, unsigned long long epoch, int logkp, int cap) {
	// chained after a Lloyd launch: let the next Lloyd launch start its prologue, then wait for the centres
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	__shared__ double c[kGridMaxK * 3], qn[kGridMaxK], cw[3], c0[3];
	__shared__ int clist[8][32];
	const int t = threadIdx.x, lane = t & 31, wib = t >> 5;
	constexpr int W = HALF ? 16 : 32;                 // lanes per cell
	const int sub = lane & (W - 1), half = HALF ? lane >> 4 : 0;
	const uint32_t hmask = HALF ? 0xFFFFu << (16 * half) : 0xffffffffu;
	if (t < kGridMaxK) {
		const bool ok = t < K;
		const double cx = ok ? centers[3 * t] : 0.0, cy = ok ? centers[3 * t + 1] : 0.0, cz = ok ? centers[3 * t + 2] : 0.0;
		c[3 * t] = cx; c[3 * t + 1] = cy; c[3 * t + 2] = cz;
		qn[t] = cx * cx + cy * cy + cz * cz;
	}
	if (t < 3) {
		// the kernel puts x into cell i when sat(x s + o) gs is in [i, i+1): cell width and origin in feature units
		cw[t] = 1.0 / ((double)g.gs[t] * (double)g.s[t]);
		c0[t] = -(double)g.o[t] / (double)g.s[t];
	}
	uint32_t *pool = out + cap;  // the pool follows the `cap` cell words of this KP (KCfg::kGridCapUsed)
	// two pool counters per build (tier 1: the crowded cells among the M cells the Lloyd kernel has marked, entries
	// [0, min(M, kGridTier1)); tier 2: the other crowded cells, the entries behind), ping-pong by epoch; block 0 clears
	// the other pair for the next build
	unsigned int *ctr = reinterpret_cast<unsigned int *>(out + kGridWords) + 2 * (epoch & 1ull);
	const uint32_t *marks = out + kGridWords + 4;
	if (blockIdx.x == 0 && t < 2) reinterpret_cast<unsigned int *>(out + kGridWords)[2 * ((epoch + 1) & 1ull) + t] = 0u;
	// (tiers only where the crowded cells outnumber the pool: K > 32 — measured at K = 32, where the 768 entries cover
	// all but ~16 of the crowded cells, reserving entries for marked cells costs more walks than it saves)
	__shared__ unsigned int s_marked;
	if (t == 0) s_marked = 0u;
	__syncthreads();
	if (logkp > 5) {
		unsigned int m = 0u;
		for (int w = t; w < kGridMarkWords; w += blockDim.x) m += (unsigned int)__popc(marks[w]);
		for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
		if (lane == 0 && m) atomicAdd(&s_marked, m);
	}
	__syncthreads();
	int *mine = clist[wib] + W * half;
	const int g01 = g.g[0] * g.g[1];
	const float inv_g01 = 1.0f / (float)g01, inv_g0 = 1.0f / (float)g.g[0];
	constexpr int kPer = HALF ? 2 : 1;  // cells a warp builds at a time
	for (int first = (blockIdx.x * 8 + wib) * kPer; first < g.ncell; first += gridDim.x * 8 * kPer) {
		// the second half of a warp may run past the last cell: it rebuilds the last cell and does not store
		const bool valid = first + half < g.ncell;
		const int cell = valid ? first + half : g.ncell - 1;
		// (cell < 2^14 and the divisors <= 2^12: the float quotient of cell + 0.5 cannot cross an integer)
		int idx[3];
		idx[2] = (int)(((float)cell + 0.5f) * inv_g01);
		const int rem = cell - idx[2] * g01;
		idx[1] = (int)(((float)rem + 0.5f) * inv_g0);
		idx[0] = rem - idx[1] * g.g[0];
		double lo[3], hi[3];
		bool lo_inf[3], hi_inf[3];
#pragma unroll
		for (int j = 0; j < 3; ++j) {
			lo[j] = c0[j] + cw[j] * ((double)idx[j] - 1.0 / 512.0);          // box inflated by 2^-9 of a cell
			hi[j] = c0[j] + cw[j] * ((double)idx[j] + 1.0 + 1.0 / 512.0);
			lo_inf[j] = idx[j] == 0; hi_inf[j] = idx[j] == g.g[j] - 1;       // border cells collect everything beyond the box
		}
		auto dominated = [&](int k, int w) {  // w closer than k everywhere in the box
			double m = qn[w] - qn[k];
			bool unbounded = false;
#pragma unroll
			for (int j = 0; j < 3; ++j) {
				const double u = c[3 * k + j] - c[3 * w + j];
				if (u > 0.0) { unbounded = unbounded || hi_inf[j]; m += 2.0 * hi[j] * u; }
				else if (u < 0.0) { unbounded = unbounded || lo_inf[j]; m += 2.0 * lo[j] * u; }
			}
			return !unbounded && m < -1e-7 * (1.0 + qn[w] + qn[k]);
		};
		// w* = centre nearest to the middle of the (finite) cell
		double bd = 1e300;
		int bw = 0x7fffffff;
		for (int k = sub; k < K; k += W) {
			double d = 0.0;
#pragma unroll
			for (int j = 0; j < 3; ++j) { const double dj = 0.5 * (lo[j] + hi[j]) - c[3 * k + j]; d += dj * dj; }
			if (d < bd) { bd = d; bw = k; }
		}
		for (int o = W / 2; o > 0; o >>= 1) {
			const double od = __shfl_xor_sync(hmask, bd, o);
			const int ow = __shfl_xor_sync(hmask, bw, o);
			if (od < bd || (od == bd && ow < bw)) { bd = od; bw = ow; }
		}
		int cnt = 0;
		for (int k0 = 0; k0 < K; k0 += W) {
			const int k = k0 + sub;
			const bool cand = k < K && (k == bw || !dominated(k, bw));
			const uint32_t m = (__ballot_sync(hmask, cand) >> (HALF ? 16 * half : 0)) & (HALF ? 0xFFFFu : 0xffffffffu);
			if (cand) {
				const int pos = cnt + __popc(m & ((1u << sub) - 1u));
				if (pos < W) mine[pos] = k;
			}
			cnt += __popc(m);
		}
		__syncwarp(hmask);
		if (cnt > 1 && cnt <= W) {  // refine: drop a candidate that another candidate dominates
			bool keep = sub < cnt;
			if (keep)
				for (int j = 0; j < cnt; ++j)
					if (j != sub && dominated(mine[sub], mine[j])) { keep = false; break; }
			const uint32_t m = (__ballot_sync(hmask, keep) >> (HALF ? 16 * half : 0)) & (HALF ? 0xFFFFu : 0xffffffffu);
			const int mylab = sub < cnt ? mine[sub] : 0;
			__syncwarp(hmask);
			if (keep) mine[__popc(m & ((1u << sub) - 1u))] = mylab;
			cnt = __popc(m);
			__syncwarp(hmask);
		}
		// Entry of the cell, computed by every lane alike on bit masks (a lane-0 section with label lists and an
		// insertion sort was a quarter of this kernel's instructions): the candidates as a 64-bit mask, padded with
		// the lowest-numbered other centres up to four / eight labels, read out in ascending order.
		const int cand = (sub < cnt && cnt <= W) ? mine[sub] : -1;
		const uint32_t cm_lo = __reduce_or_sync(hmask, (cand >= 0 && cand < 32) ? 1u << cand : 0u);
		const uint32_t cm_hi = __reduce_or_sync(hmask, cand >= 32 ? 1u << (cand - 32) : 0u);
		const unsigned long long all_k = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
		auto padded = [&](int want) {  // candidates + the lowest (want - cnt) centres that are not candidates
			unsigned long long m = ((unsigned long long)cm_hi << 32) | cm_lo, rest = all_k & ~m;
			for (int i = cnt; i < want; ++i) { const unsigned long long low = rest & (0ull - rest); m |= low; rest ^= low; }
			return m;
		};
		auto take = [](unsigned long long &m) { const int l = __ffsll((long long)m) - 1; m &= m - 1ull; return (uint32_t)l; };
		uint32_t entry;
		if (logkp <= 4) {
			// K <= 16: eight nibbles per entry, no pool (see assign_grid_nib)
			entry = 0xFFFF0000u;
			const int want = cnt <= 4 ? 4 : 8;
			if (cnt <= 8 && K >= want) {
				unsigned long long m = padded(want);
				entry = 0u;
				for (int sl = 0; sl < want; ++sl) entry |= take(m) << (4 * sl);
			}
		} else if (cnt <= 4) {
			unsigned long long m = padded(4);
			entry = 0u;
			for (int sl = 0; sl < 4; ++sl) entry |= take(m) << (8 * sl);
		} else {
			entry = 2u;  // byte0 = 2 > byte1 = 0: all K centres
			if (!HALF && cnt <= 8 && K >= 8) {
				unsigned int pi = 0u;
				if (lane == 0) {
					const bool marked = logkp > 5 && ((marks[cell >> 5] >> (cell & 31)) & 1u);
					const unsigned int tier1 = s_marked < (unsigned int)kGridTier1 ? s_marked : (unsigned int)kGridTier1;
					pi = marked ? atomicAdd(&ctr[0], 1u) : tier1 + atomicAdd(&ctr[1], 1u);
					if (marked && pi >= tier1) pi = (unsigned int)kGridPool;  // (more marked cells than tier 1 holds)
				}
				pi = __shfl_sync(0xffffffffu, pi, 0);
				if (pi < (unsigned int)kGridPool && (pi >> logkp) < (1u << logkp)) {
					unsigned long long m = padded(8);
					uint32_t w0 = 0u, w1 = 0u;
					for (int sl = 0; sl < 4; ++sl) w0 |= take(m) << (8 * sl);
					for (int sl = 0; sl < 4; ++sl) w1 |= take(m) << (8 * sl);
					if (lane == 0) { pool[2 * pi] = w0; pool[2 * pi + 1] = w1; }
					entry = 1u | ((pi & ((1u << logkp) - 1u)) << 16) | ((pi >> logkp) << 24);
				}
			}
		}
		if (sub == 0 && valid) out[cell] = entry;
		__syncwarp();
	}
}

#ifdef CS_PHASE_TIMING
#define CS_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && p.phase_ts) p.phase_ts[i] = mg_globaltimer(); } while (0)
#else
#define CS_STAMP(i) do { } while (0)
#endif

template <int KP, int FM, bool TIE, bool INERTIA, class V, bool GRID = false>
__global__ void __launch_bounds__(V::THREADS, 1) lloyd_kernel(const LloydParams p) {
	using S = Smem<KP, FM, V, GRID>;
	static_assert(!GRID || (KP >= 16 && KP <= 64 && FM == FM_F32 && TIE && !INERTIA), "GRID kernels: K <= 64, planar fp32, exact labels");
	constexpr int kPlanes = S::kPlanes, kStages = S::kStages;
	constexpr int kCopies = KCfg<KP>::kCopies;
	constexpr int kNW = V::NW, kNC = V::NC, kThreads = V::THREADS, kTile = V::TILE, U = V::U;
	extern __shared__ __align__(128) unsigned char smem[];
	float *ring = reinterpret_cast<float *>(smem);
	float4 *acc = reinterpret_cast<float4 *>(smem + S::kOffAcc);
	float4 *tab = reinterpret_cast<float4 *>(smem + S::kOffTab);
	double *c64 = reinterpret_cast<double *>(smem + S::kOffC64);
	double *red = reinterpret_cast<double *>(smem + S::kOffRed);
	float *lut = reinterpret_cast<float *>(smem + S::kOffLut);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::kOffBar);
	uint64_t *empty = full + kStages;
	uint64_t *gridbar = empty + kStages;  // GRID: completion of the candidate-table copy
	__shared__ int s_is_last;
	__shared__ int s_lost;  // a partial word never arrived (cannot happen; bounds the poll): totals become NaN

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int K = p.K;
	const long long n = p.n;
	// batched launch: blockIdx.y selects the image; every per-image array is offset here (0 when not batched)
	const long long img = blockIdx.y;
	const float *f0 = p.f0 + img * p.img_stride_px, *f1 = p.f1 + img * p.img_stride_px, *f2 = p.f2 + img * p.img_stride_px;
	const uint32_t *rgba = p.rgba + img * p.img_stride_px;
	const double *centers_in = p.centers + img * (K * 3);
	uint8_t *labels = p.labels ? p.labels + img * p.img_stride_label : nullptr;
	unsigned long long *pwords = p.pwords + (size_t)img * gridDim.x * (2 * kMaxPartialVals);
	const uint32_t ltag = (uint32_t)p.launch_epoch;
	unsigned int *counter = p.counter + img;
	const long long ntiles = (n + kTile - 1) / kTile;

	// ---- prologue, part A: barriers, zero accumulators, first ring loads ----
	// Nothing in part A reads what an earlier Lloyd launch on the stream wrote (the pixels and the feature
	// table are inputs), so in a CS_LLOYD_CHAINED launch (programmatic dependent launch) it runs while the
	// previous iteration's last CTA is still combining; everything after griddepcontrol.wait sees that
	// launch's centres and control block.
	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kNW); }
		if (GRID) mbar_init(gridbar, 1);
		mbar_fence_init();
		s_lost = 0;
	}
	if (FM == FM_RGBA8)
		for (int i = tid; i < 3 * 256; i += kThreads) lut[i] = p.lut3 ? p.lut3[i] : (float)(i & 255);
	for (int i = tid; i < S::kAccBytes / 16; i += kThreads) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	__syncthreads();
	// producer lane: request tile `it` of this CTA into ring stage it % kStages
	auto issue_tile = [&](int it, long long tile) {
		const int s = it % kStages;
		const long long base = tile * (long long)kTile;
		long long rem = n - base;
		if (rem > kTile) rem = kTile;
		const uint32_t bytes = (uint32_t)(rem & ~3LL) * 4u;  // whole 16-byte groups only
		mbar_arrive_expect_tx(&full[s], bytes * kPlanes);
		if (bytes) {
			float *dst = ring + (size_t)s * kPlanes * kTile;
			if (FM == FM_F32) {
				bulk_g2s(dst, f0 + base, bytes, &full[s]);
				bulk_g2s(dst + kTile, f1 + base, bytes, &full[s]);
				bulk_g2s(dst + 2 * kTile, f2 + base, bytes, &full[s]);
			} else {
				bulk_g2s(dst, rgba + base, bytes, &full[s]);
			}
		}
	};
	// the barriers are initialised and visible: start the first kStages loads now, so that their HBM
	// latency overlaps the rest of the prologue (key constants, centre table)
	if (tid == kNC) {
		int it = 0;
		for (long long tile = blockIdx.x; tile < ntiles && it < kStages; tile += gridDim.x, ++it) issue_tile(it, tile);
	}
	// ---- part B: after the previous launch of the stream has completed and flushed ----
	asm volatile("griddepcontrol.wait;" ::: "memory");
	// the next chained launch may be scheduled from now on: its CTAs take an SM as ours retire and wait
	// at the line above for this grid to finish
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	if (p.ctl && *reinterpret_cast<const volatile double *>(p.ctl) != 0.0) {
		// halted by an earlier launch: let the loads requested above land before the CTA retires
		int it = 0;
		for (long long tile = blockIdx.x; tile < ntiles && it < kStages; tile += gridDim.x, ++it) mbar_wait(&full[it], 0);
		return;
	}
	CS_STAMP(0);
	if (GRID && tid == kNC) {
		// the candidate table grid_build_kernel wrote for these centres (it precedes this launch on the stream)
		mbar_arrive_expect_tx(gridbar, (uint32_t)S::kGridBytes);
		bulk_g2s(smem + S::kOffGrid, p.grid_tab, (uint32_t)S::kGridBytes, gridbar);
	}
	for (int i = tid; i < KP * 3; i += kThreads) c64[i] = (i < K * 3) ? centers_in[i] : 0.0;
	__syncthreads();
	__shared__ KeyConst s_kc;
	constexpr bool INTKEY = KP <= 64;
	if (tid == 0) {
		double m = 0.0;
		for (int k = 0; k < K; ++k) {
			double cx = c64[3 * k], cy = c64[3 * k + 1], cz = c64[3 * k + 2];
			m = fmax(m, cx * cx + cy * cy + cz * cz);
		}
		// integer keys: v = (|c|^2 - 2 x.c + x2max) S + 1 must stay inside the 2^(32-b) window
		// of bit patterns above bits(1.0f), i.e. below 2^(2^(9-b)) — S is the largest power of
		// two <= 1 that keeps the top value under half of that.
		const double vtop = (sqrt(p.x2max) + sqrt(m)) * (sqrt(p.x2max) + sqrt(m));
		double S = 1.0;
		if (INTKEY) {
			const int binades = 1 << (9 - KCfg<KP>::kBits);
			const double lim = binades >= 64 ? 9.0e18 : (double)(1ull << (binades - 1));
			while (vtop * S + 1.0 >= lim) S *= 0.5;
		}
		if (GRID) {
			// fixed-point keys: every partial sum of (|c|^2 - 2 x.c + x2max) s stays inside [0, 0.9 * 2^22]
			const double vmax = (sqrt(p.x2max) + sqrt(m)) * (sqrt(p.x2max) + sqrt(m)) + p.x2max;
			S = 0.9 * 4194304.0 / (vmax > 1e-30 ? vmax : 1e-30);
		}
		s_kc.S = (float)S;
		s_kc.x2max = (float)p.x2max;
		// 3 FMA roundings + 4 rounded table entries, each <= 2^-24 of the largest partial sum
		// (<= vtop S + 1), for two keys, with 1.5x slack
		s_kc.tau_v = (float)(1.5 * 2.0 * 7.0 * 5.9604645e-8 * (vtop * S + 1.0));
		s_kc.cnmax = (float)m;
		s_kc.sh = 1u << KCfg<KP>::kBits;
		{
			const int binades = 1 << (9 - KCfg<KP>::kBits);
			s_kc.pad = binades >= 64 ? 1.0e19f : (float)((double)(1ull << (binades >= 63 ? 62 : binades)) * 0.99);
		}
	}
	__syncthreads();
	if (GRID) {
		// 16-byte entries {-2s cx, -2s cy, -2s cz, (|c|^2 + x2max) s + 1.5 * 2^23}, kTabCopies consecutive copies each
		if (tid < KP * S::kTabCopies) {
			const int k = tid / S::kTabCopies;
			const double S = (double)s_kc.S;
			float4 v = make_float4(0.f, 0.f, 0.f, 16777215.0f);  // padding entry: the largest key
			if (k < K) {
				const double cx = c64[3 * k], cy = c64[3 * k + 1], cz = c64[3 * k + 2];
				v = make_float4((float)(-2.0 * cx * S), (float)(-2.0 * cy * S), (float)(-2.0 * cz * S),
				                (float)((cx * cx + cy * cy + cz * cz + p.x2max) * S + (double)kGridMagic));
			}
			tab[tid] = v;
		}
	} else if (tid < KP / 2) {
		// pair table: {-2cx_k, -2cx_k+1, -2cy_k, -2cy_k+1}, {-2cz_k, -2cz_k+1, q_k, q_k+1},
		// q = |c|^2 (float keys) or (|c|^2 + x2max) S + 1 with the -2c terms scaled by S (integer keys)
		const double S = INTKEY ? (double)s_kc.S : 1.0;
		float v[2][4];
		for (int h = 0; h < 2; ++h) {
			int k = 2 * tid + h;
			if (k < K) {
				double cx = c64[3 * k], cy = c64[3 * k + 1], cz = c64[3 * k + 2];
				const double cn = cx * cx + cy * cy + cz * cz;
				v[h][0] = (float)(-2.0 * cx * S); v[h][1] = (float)(-2.0 * cy * S);
				v[h][2] = (float)(-2.0 * cz * S);
				v[h][3] = INTKEY ? (float)((cn + p.x2max) * S + 1.0) : (float)cn;
			} else {  // padding entry: never the minimum
				v[h][0] = v[h][1] = v[h][2] = 0.f; v[h][3] = INTKEY ? s_kc.pad : 1.0e30f;
			}
		}
		tab[2 * tid] = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
		tab[2 * tid + 1] = make_float4(v[0][2], v[1][2], v[0][3], v[1][3]);
	}
	__syncthreads();

	CS_STAMP(1);
	if (warp == kNW) {
		// ================= producer warp =================
		// (the first kStages tiles were already requested in the prologue)
		if (lane == 0) {
			int it = 0;
			for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
				if (it < kStages) continue;
				mbar_wait_backoff(&empty[it % kStages], ((it / kStages) - 1) & 1);
				issue_tile(it, tile);
			}
		}
	} else {
		// ================= consumer warps =================
		float4 *wacc = acc + (size_t)warp * KP * kCopies;
		const uint32_t keymask = p.keymask;
		const KeyConst kc = s_kc;
		float2 treg[KP <= 16 ? KP * 2 : 1];
		if (V::TABREG && KP <= 16) {
#pragma unroll
			for (int pr = 0; pr < KP / 2; ++pr) {
				const float4 t0 = tab[2 * pr], t1 = tab[2 * pr + 1];
				treg[4 * pr] = make_float2(t0.x, t0.y); treg[4 * pr + 1] = make_float2(t0.z, t0.w);
				treg[4 * pr + 2] = make_float2(t1.x, t1.y); treg[4 * pr + 3] = make_float2(t1.z, t1.w);
			}
		}
		float inert = 0.f;
		double inert64 = 0.0;
		const uint32_t tab_s = smem_u32(tab);
		GridConst gc;
		if (GRID) {
			gc.sx = p.grid.s[0]; gc.ox = p.grid.o[0]; gc.gx = p.grid.gs[0];
			gc.sy = p.grid.s[1]; gc.oy = p.grid.o[1]; gc.gy = p.grid.gs[1];
			gc.sz = p.grid.s[2]; gc.oz = p.grid.o[2]; gc.gz = p.grid.gs[2];
			gc.stride_y = (uint32_t)p.grid.g[0]; gc.stride_z = (uint32_t)(p.grid.g[0] * p.grid.g[1]);
			gc.base_c = smem_u32(smem + S::kOffGrid) - 4u * __float_as_uint(kGridMagic) * (1u + gc.stride_y + gc.stride_z);
			gc.pool_s = smem_u32(smem + S::kOffGrid) + (uint32_t)(KCfg<KP>::kGridCapUsed * 4);
			gc.ctab_s = tab_s + (uint32_t)((lane & (S::kTabCopies - 1)) << 4);  // the lane's copy of the table
			gc.logkp = (uint32_t)KCfg<KP>::kBits;
			gc.marks = KP > 32 ? p.grid_marks : nullptr;
			mbar_wait(gridbar, 0);
		}
		bool ready = false;
		int it = 0;
		for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
			const int s = it % kStages;
			const long long base = tile * (long long)kTile;
			long long rem = n - base;
			if (rem > kTile) rem = kTile;
			constexpr int P = 4 * U;
			float x[P], y[P], z[P];
			bool use[P];
			int lab[P];
			if (!ready) mbar_wait(&full[s], (it / kStages) & 1);
			const float *stage = ring + (size_t)s * kPlanes * kTile;
			if (FM == FM_F32 && rem == kTile) {
				// ---- complete tile of fp32 planes (all but the image's last tile): no validity logic ----
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int px0 = (u * kNC + tid) * 4;
					const float4 a = *reinterpret_cast<const float4 *>(stage + px0);
					const float4 b = *reinterpret_cast<const float4 *>(stage + kTile + px0);
					const float4 c = *reinterpret_cast<const float4 *>(stage + 2 * kTile + px0);
					x[4 * u] = a.x; x[4 * u + 1] = a.y; x[4 * u + 2] = a.z; x[4 * u + 3] = a.w;
					y[4 * u] = b.x; y[4 * u + 1] = b.y; y[4 * u + 2] = b.z; y[4 * u + 3] = b.w;
					z[4 * u] = c.x; z[4 * u + 1] = c.y; z[4 * u + 2] = c.z; z[4 * u + 3] = c.w;
				}
#pragma unroll
				for (int q = 0; q < P; ++q) use[q] = true;
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty[s]);
				// a non-blocking look at the NEXT tile's barrier now: when the data is already there (the usual
				// case) the next iteration starts without waiting for a barrier query's round trip
				ready = (tile + gridDim.x < ntiles) && mbar_test(&full[(it + 1) % kStages], ((it + 1) / kStages) & 1);
				if constexpr (GRID) {
					if constexpr (KP <= 16) assign_grid_nib<true, P, KCfg<KP>::kGridTabShift>(x, y, z, use, lab, gc, c64, K);
					else assign_grid<true, P, KCfg<KP>::kGridTabShift, (KP > 32)>(x, y, z, use, lab, gc, c64, K);
					update_slots<KP, true, P>(x, y, z, use, lab, wacc, lane);
				} else {
					assign_update<KP, FM, TIE, INERTIA, V, true, P>(x, y, z, use, lab, tab_s, treg, c64, K, keymask, kc, wacc, lane, inert);
				}
				if (INERTIA) { inert64 += (double)inert; inert = 0.f; }
				if (labels) {
#pragma unroll
					for (int u = 0; u < U; ++u) {
						// four label bytes -> one word with three byte permutes
						const uint32_t lo = __byte_perm((uint32_t)lab[4 * u], (uint32_t)lab[4 * u + 1], 0x1140);
						const uint32_t hi = __byte_perm((uint32_t)lab[4 * u + 2], (uint32_t)lab[4 * u + 3], 0x1140);
						*reinterpret_cast<uint32_t *>(labels + base + (u * kNC + tid) * 4) = __byte_perm(lo, hi, 0x5410);
					}
				}
				continue;
			}
			bool all_use = true;
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int px0 = (u * kNC + tid) * 4;
				int valid = (int)rem - px0;
				valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
				uint32_t raw[4];
				if (valid == 4) {
					if (FM == FM_F32) {
						const float4 a = *reinterpret_cast<const float4 *>(stage + px0);
						const float4 b = *reinterpret_cast<const float4 *>(stage + kTile + px0);
						const float4 c = *reinterpret_cast<const float4 *>(stage + 2 * kTile + px0);
						x[4 * u] = a.x; x[4 * u + 1] = a.y; x[4 * u + 2] = a.z; x[4 * u + 3] = a.w;
						y[4 * u] = b.x; y[4 * u + 1] = b.y; y[4 * u + 2] = b.z; y[4 * u + 3] = b.w;
						z[4 * u] = c.x; z[4 * u + 1] = c.y; z[4 * u + 2] = c.z; z[4 * u + 3] = c.w;
					} else {
						const uint4 a = *reinterpret_cast<const uint4 *>(stage + px0);
						raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
					}
				} else {
					// ragged tail (< 4 valid pixels): these were not part of the bulk copy
#pragma unroll
					for (int q = 0; q < 4; ++q) {
						const bool ok = q < valid;
						if (FM == FM_F32) {
							x[4 * u + q] = ok ? f0[base + px0 + q] : 0.f;
							y[4 * u + q] = ok ? f1[base + px0 + q] : 0.f;
							z[4 * u + q] = ok ? f2[base + px0 + q] : 0.f;
						} else {
							raw[q] = ok ? rgba[base + px0 + q] : 0u;
						}
					}
				}
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					bool ok = q < valid;
					if (FM == FM_RGBA8) {
						const uint32_t r = raw[q] & 0xFFu, g = (raw[q] >> 8) & 0xFFu,
						               b = (raw[q] >> 16) & 0xFFu, a = raw[q] >> 24;
						x[4 * u + q] = lut[r]; y[4 * u + q] = lut[256 + g]; z[4 * u + q] = lut[512 + b];
						const int bright = p.mask_mode == 0 ? (int)(r + g + b) : (int)b;
						ok = ok && a > 0u && bright > p.min_rgb_sum;
					}
					use[4 * u + q] = ok;
					all_use = all_use && ok;
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[s]);
			ready = (tile + gridDim.x < ntiles) && mbar_test(&full[(it + 1) % kStages], ((it + 1) / kStages) & 1);

			// warp-uniform fast path when every pixel of the warp's groups is real and unmasked
			if constexpr (GRID) {
				if constexpr (KP <= 16) assign_grid_nib<false, P, KCfg<KP>::kGridTabShift>(x, y, z, use, lab, gc, c64, K);
				else assign_grid<false, P, KCfg<KP>::kGridTabShift, (KP > 32)>(x, y, z, use, lab, gc, c64, K);
				update_slots<KP, false, P>(x, y, z, use, lab, wacc, lane);
			} else if (__all_sync(0xffffffffu, all_use))
				assign_update<KP, FM, TIE, INERTIA, V, true, P>(x, y, z, use, lab, tab_s, treg, c64, K, keymask, kc, wacc, lane, inert);
			else
				assign_update<KP, FM, TIE, INERTIA, V, false, P>(x, y, z, use, lab, tab_s, treg, c64, K, keymask, kc, wacc, lane, inert);
			if (INERTIA) { inert64 += (double)inert; inert = 0.f; }

			// ---- labels: one 32-bit word per 4 pixels ----
			if (labels) {
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int px0 = (u * kNC + tid) * 4;
					const int valid = (int)rem - px0;
					uint8_t *lp = labels + base + px0;
					if (valid >= 4) {
						uint32_t w = 0;
#pragma unroll
						for (int q = 0; q < 4; ++q) w |= (uint32_t)(use[4 * u + q] ? lab[4 * u + q] : 255) << (8 * q);
						*reinterpret_cast<uint32_t *>(lp) = w;
					} else {
#pragma unroll
						for (int q = 0; q < 4; ++q)
							if (q < valid) lp[q] = (uint8_t)(use[4 * u + q] ? lab[4 * u + q] : 255);
					}
				}
			}
		}
		if (INERTIA) {
			// warp-reduce the per-thread fp64 inertia (fixed butterfly), one slot per warp
			for (int o = 16; o > 0; o >>= 1) inert64 += __shfl_xor_sync(0xffffffffu, inert64, o);
			if (lane == 0) red[KP * 4 + warp] = inert64;
		}
	}
	__syncthreads();

	CS_STAMP(2);
	// ---- CTA epilogue: fold the lane-private fp32 slots to fp64, fixed order ----
	// A warp's slot region is rows of 32 consecutive float4 (= the kCopies copies of 32 / kCopies adjacent
	// clusters).  One warp takes one row: lane l reads float4 l of that row in every consumer warp's region
	// (512 contiguous bytes per load: conflict-free), adds the kNW values in warp order in fp64, and a
	// butterfly over the kCopies lanes of each cluster finishes the sum.  output o = k*4 + c (c = 3: count).
	{
		constexpr int kOut = KP * 4;
		constexpr int kRows = KP * kCopies / 32;
		unsigned long long *mine = pwords + (size_t)blockIdx.x * (2 * kMaxPartialVals);
		for (int row = warp; row < kRows; row += kThreads / 32) {
			double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
			for (int w = 0; w < kNW; ++w) {
				const float4 v = acc[(size_t)w * KP * kCopies + row * 32 + lane];
				s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
			}
#pragma unroll
			for (int o = kCopies / 2; o > 0; o >>= 1) {
				s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
				s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
			}
			if ((lane % kCopies) == 0) {
				unsigned long long *d = mine + (size_t)(row * (32 / kCopies) + lane / kCopies) * 8;  // output o -> words 2o, 2o+1
				put_double_gpu(d, s0, ltag); put_double_gpu(d + 2, s1, ltag);
				put_double_gpu(d + 4, s2, ltag); put_double_gpu(d + 6, s3, ltag);
			}
		}
		if (INERTIA && tid == 0) {
			double s = 0.0;
			for (int w = 0; w < kNW; ++w) s += red[KP * 4 + w];
			put_double_gpu(mine + 2 * kOut, s, ltag);
		}
	}

	// ---- last CTA: global combine in block order (+ fused M-step tail) ----
	CS_STAMP(3);
	// No fence: the partials are self-validating words (tag = this launch's epoch), so the counter only elects
	// the CTA that combines; it then reads every CTA's words and, should one not have landed yet, polls it.
	// (A gpu-scope fence here made every CTA wait for its own label stores; with peer access enabled — the
	// multi-GPU runs — that wait grew with the number of peers.)
	__syncthreads();
	if (tid == 0) {
		const unsigned int prev = atomicAdd(counter, 1u);
		s_is_last = (prev == gridDim.x - 1);
	}
	__syncthreads();
	if (!s_is_last) return;
	CS_STAMP(4);
	// control block: fetched now so that the round trip overlaps the combine (only this thread writes it)
	double ctl_iters = 0.0, ctl_tol = 0.0;
	if (p.ctl && p.centers_out && tid == 0) { ctl_iters = p.ctl[1]; ctl_tol = p.ctl[2]; }
	double *out_sums = p.sums + img * (K * 3), *out_counts = p.counts + img * K;
	double *out_inertia = p.inertia ? p.inertia + img : nullptr;
	// shared mirrors of the totals for the fused M-step tail (ring space past the combine scratch)
	static_assert(S::kRingBytes >= 16384 + KP * 4 * 8, "tail mirrors");
	double *fin_sums = reinterpret_cast<double *>(smem + 16384), *fin_counts = fin_sums + KP * 3;
	{
		constexpr int kOut = KP * 4;
		constexpr int kVals = kOut + (INERTIA ? 1 : 0);
		const bool mg = p.world > 1;
		const int par = (int)(p.epoch & 1ull);
		// The per-CTA partials are summed in a FIXED order (run-to-run deterministic): kSplit contiguous
		// block ranges are summed in block order by different threads (so the L2 round trips overlap),
		// then the kSplit range sums are added in range order.
		constexpr int kSplit = (kThreads / kVals) > 16 ? 16 : ((kThreads / kVals) > 1 ? (kThreads / kVals) : 1);
		double *scratch2 = reinterpret_cast<double *>(smem);  // the ring is idle now
		if (kSplit > 1) {
			for (int item = tid; item < kVals * kSplit; item += kThreads) {
				const int o = item % kVals, part = item / kVals;
				const unsigned int b0 = (unsigned int)part * gridDim.x / kSplit, b1 = (unsigned int)(part + 1) * gridDim.x / kSplit;
				double s = 0.0;
				for (unsigned int b = b0; b < b1; ++b) s += get_double_gpu(pwords + ((size_t)b * kMaxPartialVals + o) * 2, ltag, &s_lost);
				scratch2[part * kVals + o] = s;
			}
			__syncthreads();
		}
		for (int o = tid; o < kVals; o += kThreads) {
			double s = 0.0;
			if (kSplit > 1) {
				for (int part = 0; part < kSplit; ++part) s += scratch2[part * kVals + o];
			} else {
				for (unsigned int b = 0; b < gridDim.x; ++b)
					s += get_double_gpu(pwords + ((size_t)b * kMaxPartialVals + o) * 2, ltag, &s_lost);
			}
			if (mg) {
				// push this rank's partial into slot [par][rank] of every rank's mailbox (NVLink P2P stores):
				// two self-validating 8-byte words per double
				const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
				const uint32_t tag = (uint32_t)p.epoch;
				for (int q = 0; q < p.world; ++q) {
					unsigned long long *w = &p.mb[q]->word[par][p.rank][2 * o];
					word_store_sys(w, (uint32_t)bits, tag);
					word_store_sys(w + 1, (uint32_t)(bits >> 32), tag);
				}
			} else if (o == kOut) {
				if (out_inertia) *out_inertia = s;
			} else {
				const int k = o >> 2, c = o & 3;
				if (k < K) {
					if (c == 3) { out_counts[k] = s; fin_counts[k] = s; } else { out_sums[3 * k + c] = s; fin_sums[3 * k + c] = s; }
				}
			}
		}
		if (tid == 0) *counter = 0u;  // re-arm for the next launch on this stream
		if (mg) {
			__shared__ int s_timeout;
			if (tid == 0) s_timeout = 0;
			__syncthreads();
			// every value thread collects its value from all ranks' slots of OUR mailbox (bounded spin on the
			// tags) and adds them in rank order: bit-identical totals on every GPU
			const cs_mailbox *own = p.mb[p.rank];
			const uint32_t tag = (uint32_t)p.epoch;
			for (int o = tid; o < kVals; o += kThreads) {
				double s = 0.0;
				bool bad = false;
				const unsigned long long t0 = mg_globaltimer();
				for (int q = 0; q < p.world; ++q) {
					const unsigned long long *w = &own->word[par][q][2 * o];
					unsigned long long lo = word_load_sys(w), hi = word_load_sys(w + 1);
					while ((uint32_t)(lo >> 32) != tag || (uint32_t)(hi >> 32) != tag) {
						if (mg_globaltimer() - t0 > 20000000000ull) { bad = true; break; }
						__nanosleep(20);
						lo = word_load_sys(w); hi = word_load_sys(w + 1);
					}
					s += __longlong_as_double((long long)((hi << 32) | (lo & 0xFFFFFFFFull)));
				}
				if (bad) { s = __longlong_as_double(0x7ff8000000000000ll); s_timeout = 1; }
				if (o == kOut) {
					if (out_inertia) *out_inertia = s;
				} else {
					const int k = o >> 2, c = o & 3;
					if (k < K) {
						if (c == 3) { out_counts[k] = s; fin_counts[k] = s; } else { out_sums[3 * k + c] = s; fin_sums[3 * k + c] = s; }
					}
				}
			}
			__syncthreads();
			if (s_timeout != 0 && tid == 0) p.mb[p.rank]->error = p.epoch;
		}
	}
	CS_STAMP(5);
	if (p.centers_out) {
		__syncthreads();
		double shift2_total = 0.0;
		int n_empty = 0;
		finalize_block(fin_sums, fin_counts, c64, K, p.centers_out + img * (K * 3), p.stats + img * 4, red,
		               shift2_total, n_empty);
		if (p.ctl && tid == 0) {
			if (n_empty > 0) {
				p.ctl[0] = 2.0;  // an empty cluster: this iteration has to be redone with relocation by the host
			} else {
				p.ctl[1] = ctl_iters + 1.0;
				if (shift2_total <= ctl_tol) p.ctl[0] = 1.0;  // sum of squared shifts <= tol: converged
			}
		}
		CS_STAMP(6);
	}
}

__global__ void __launch_bounds__(256, 1)
finalize_kernel(const double *sums, const double *counts, const double *c_old, int K,
                double *c_new, double *stats) {
	__shared__ double shift2[CS_MAX_K];
	double sh2;
	int ne;
	finalize_block(sums, counts, c_old, K, c_new, stats, shift2, sh2, ne);
}

template <int KP, int FM, bool TIE, bool INERTIA, class V, bool GRID = false>
int launch_one(const cs_ctx *ctx, const LloydParams &p, bool chained, cudaStream_t st) {
	using S = Smem<KP, FM, V, GRID>;
	auto kern = lloyd_kernel<KP, FM, TIE, INERTIA, V, GRID>;
	static bool attr_done[16] = {};
	if (!attr_done[ctx->device & 15]) {
		CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
		attr_done[ctx->device & 15] = true;
	}
	const long long ntiles = (p.n + V::TILE - 1) / V::TILE;
	int grid = (int)(ntiles < ctx->sm_count ? (ntiles < 1 ? 1 : ntiles) : ctx->sm_count);
	int images = 1;
	if (ctx->launch_images > 1) {
		// batched: a few CTAs per image so that all images of the launch fill the SMs for several waves
		const int per = ctx->launch_ctas_per_image;
		grid = (int)(ntiles < per ? (ntiles < 1 ? 1 : ntiles) : per);
		images = ctx->launch_images;
	}
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid, images);
	cfg.blockDim = dim3(V::THREADS);
	cfg.dynamicSmemBytes = S::kTotal;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	static const bool no_pdl = getenv("CS_NO_PDL") != nullptr;  // development switch
	if (chained && !no_pdl) {
		// programmatic dependent launch: this grid may be scheduled while the previous kernel of the stream
		// (a Lloyd launch, which executes griddepcontrol.launch_dependents) is still running; the kernel
		// orders itself behind it with griddepcontrol.wait
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = 1;
	}
	CS_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
	return 0;
}

template <int KP, int FM, class V>
int launch_flags(const cs_ctx *ctx, const LloydParams &p, int flags, cudaStream_t st) {
	const bool tie = (flags & CS_LLOYD_EXACT_TIES) != 0, inert = p.inertia != nullptr;
	const bool ch = (flags & CS_LLOYD_CHAINED) != 0;
	if (tie) return inert ? launch_one<KP, FM, true, true, V>(ctx, p, ch, st) : launch_one<KP, FM, true, false, V>(ctx, p, ch, st);
	return inert ? launch_one<KP, FM, false, true, V>(ctx, p, ch, st) : launch_one<KP, FM, false, false, V>(ctx, p, ch, st);
}

// ---- grid-filtered exact assignment: geometry from the caller's feature box, table build, launch ----
#ifndef CS_GRID_NW
#define CS_GRID_NW 16  // consumer warps of the GRID kernels (development: fewer warps = smaller accumulator block = deeper ring)
#endif
using VarGrid = Var<CS_GRID_NW, 1, false, 0, 1>;  // 4 pixels per thread per tile: 2 x 24 KB ring beside the 46 KB table
constexpr long long kGridMinPixels = 1 << 18;  // below this the table build is not worth its ~3 us

GridGeom make_grid_geom(const cs_ctx *ctx, int cap) {
	GridGeom g{};
	double ext[3], vol = 1.0;
	int free_dims = 0;
	for (int j = 0; j < 3; ++j) {
		ext[j] = ctx->box_hi[j] - ctx->box_lo[j];
		if (ext[j] > 0.0) { vol *= ext[j]; ++free_dims; }
	}
	// near-cubic cells: g_j proportional to the extent, product <= cap, each <= 64
	const double cellw = free_dims ? pow(vol / (double)cap, 1.0 / free_dims) : 1.0;
	for (int j = 0; j < 3; ++j) {
		int gj = ext[j] > 0.0 ? (int)floor(ext[j] / cellw) : 1;
		g.g[j] = gj < 1 ? 1 : (gj > 64 ? 64 : gj);
	}
	while ((long long)g.g[0] * g.g[1] * g.g[2] > cap) {
		int big = 0;
		for (int j = 1; j < 3; ++j)
			if (g.g[j] > g.g[big]) big = j;
		--g.g[big];
	}
	for (int j = 0; j < 3; ++j) {
		g.s[j] = ext[j] > 0.0 ? (float)(1.0 / ext[j]) : 0.f;
		g.o[j] = ext[j] > 0.0 ? (float)(-ctx->box_lo[j] / ext[j]) : 0.f;
		g.gs[j] = (float)((double)g.g[j] * (1.0 - 1.0 / 4096.0));  // sat(..) = 1 still lands in the last cell
	}
	g.ncell = g.g[0] * g.g[1] * g.g[2];
	return g;
}

// build the candidate table for p.centers, then the GRID Lloyd launch behind it
template <int KP>
int launch_grid(cs_ctx *ctx, LloydParams &p, bool chained, cudaStream_t st) {
	p.grid = make_grid_geom(ctx, KCfg<KP>::kGridCapUsed);
	p.grid_tab = ctx->d_grid;
	p.grid_marks = ctx->d_grid + kGridWords + 4;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(ctx->sm_count * 4);
	cfg.blockDim = dim3(256);
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	static const bool no_pdl = getenv("CS_NO_PDL") != nullptr;
	if (chained && !no_pdl) {
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr;
		cfg.numAttrs = 1;
	}
	// a new fit (an unchained launch) starts with no marked cells
	if (!chained) CS_CUDA(cudaMemsetAsync(p.grid_marks, 0, sizeof(uint32_t) * kGridMarkWords, st));
	const unsigned long long epoch = ++ctx->grid_epoch;
	CS_CUDA(cudaLaunchKernelEx(&cfg, grid_build_kernel<(KP <= 16)>, p.centers, p.K, p.grid, ctx->d_grid, epoch, (int)KCfg<KP>::kBits,
	                           (int)KCfg<KP>::kGridCapUsed));
	// the build kernel executes griddepcontrol.launch_dependents at once: the Lloyd launch is always chained to it
	return launch_one<KP, FM_F32, true, false, VarGrid, true>(ctx, p, true, st);
}

bool grid_eligible(const cs_ctx *ctx, const LloydParams &p, int flags) {
	static const bool off = getenv("CS_NO_GRID") != nullptr;  // development switch: the full walk
	// default policy: only where it beats the full walk — measured on the B200 at 64 MP (tools/grid_k_sweep.py):
	// K = 6 / 8: 0.240 / 0.250 ms against 0.215 / 0.214; K = 10 / 12 / 14 / 16: 0.269 / 0.284 / 0.304 / 0.308 against
	// 0.319 / 0.320 / 0.321 / 0.322; K = 32 / 64: 0.441 / 0.747 against 0.649 / 1.258
	static const int auto_min_k = getenv("CS_GRID_MIN_K") ? atoi(getenv("CS_GRID_MIN_K")) : 9;  // (development override)
	const int kmin = ctx->grid_policy > 0 ? 4 : auto_min_k;
	// ... and only on shards large enough to pay for the table build + table load per iteration (measured after the
	// rare-path rewrite, tools/grid_size_sweep.py, grid / walk time: K = 16: 1.13 at 4 MP, 1.00 at 8.4 MP, 0.92 at
	// 16 MP, 0.88 at 32 MP; K = 32: 1.09 at 2 MP, 0.90 at 4 MP; K = 64: 1.06 at 1 MP, 0.92 at 2 MP)
	const long long nmin = ctx->grid_policy > 0 ? kGridMinPixels : (p.K <= 16 ? 10000000LL : p.K <= 32 ? (1LL << 22) : (1LL << 21));
	return !off && ctx->box_set && ctx->grid_policy >= 0 && (flags & CS_LLOYD_EXACT_TIES) && p.inertia == nullptr &&
	       ctx->launch_images <= 1 && p.K >= kmin && p.K <= kGridMaxK && p.n >= nmin && ((flags >> 8) & 15) == 0;
}

// Tuning variants (flags bits 8..11), K <= 16 planar-fp32 only; 0 = production shape.
template <int KP>
int launch_variant(const cs_ctx *ctx, const LloydParams &p, int flags, cudaStream_t st) {
#ifdef CS_TUNING_VARIANTS
	const bool tie = (flags & CS_LLOYD_EXACT_TIES) != 0;
#endif
	switch ((flags >> 8) & 15) {
#ifdef CS_TUNING_VARIANTS
#define CS_VARIANT(id, ...)                                                                         \
	case id:                                                                                        \
		return tie ? launch_one<KP, FM_F32, true, false, __VA_ARGS__>(ctx, p, false, st)             \
		           : launch_one<KP, FM_F32, false, false, __VA_ARGS__>(ctx, p, false, st);
		CS_VARIANT(1, Var<16, 2, true>)
		CS_VARIANT(2, Var<16, 2, false, 4, 0>)
		CS_VARIANT(3, Var<16, 2, false, 8, 0>)
		CS_VARIANT(4, Var<16, 2, false, 2, 1>)
		CS_VARIANT(5, Var<16, 2, false, 8, 1>)
		CS_VARIANT(6, Var<16, 2, true, 4, 1>)
		CS_VARIANT(7, Var<16, 2, true, 2, 1>)
		CS_VARIANT(8, Var<16, 2, false, 2, 0>)
		CS_VARIANT(9, Var<16, 1, false, 4, 1>)
		CS_VARIANT(10, Var<16, 1, true, 4, 1>)
		CS_VARIANT(11, Var<16, 2, true, 8, 1>)
#undef CS_VARIANT
#endif
	default:
		return launch_flags<KP, FM_F32, VarSmallK>(ctx, p, flags, st);
	}
}

template <int FM>
int launch_k(const cs_ctx *ctx, LloydParams &p, int flags, cudaStream_t st) {
	const int K = p.K;
	int kp = 8;
	while (kp < K) kp <<= 1;
	p.keymask = ~(uint32_t)(kp - 1);
	p.pwords = ctx->d_partial_words;
	p.launch_epoch = ++ctx->lloyd_epoch;  // tag of this launch's per-CTA partial words (starts at 1; the buffer at 0)
	if (FM == FM_F32 && grid_eligible(ctx, p, flags)) {
		cs_ctx *c = const_cast<cs_ctx *>(ctx);
		const bool ch = (flags & CS_LLOYD_CHAINED) != 0;
		return kp <= 16 ? launch_grid<16>(c, p, ch, st) : kp == 32 ? launch_grid<32>(c, p, ch, st) : launch_grid<64>(c, p, ch, st);
	}
	if (FM == FM_F32 && kp == 16 && p.inertia == nullptr) return launch_variant<16>(ctx, p, flags, st);
	switch (kp) {
	case 8: return launch_flags<8, FM, VarSmallK>(ctx, p, flags, st);
	case 16: return launch_flags<16, FM, VarSmallK>(ctx, p, flags, st);
	case 32: return launch_flags<32, FM, VarLargeK>(ctx, p, flags, st);
	case 64: return launch_flags<64, FM, VarLargeK>(ctx, p, flags, st);
	case 128: return launch_flags<128, FM, VarLargeK>(ctx, p, flags, st);
	default: return launch_flags<256, FM, VarLargeK>(ctx, p, flags, st);
	}
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace
} // namespace cs

using namespace cs;

extern "C" int cs_lloyd_iter_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1,
                                 const float *d_f2, int64_t n, const double *d_centers_in, int K,
                                 uint8_t *d_labels, double *d_sums, double *d_counts,
                                 double *d_centers_out, double *d_stats, double feat_norm2_max, int flags,
                                 void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_centers_in && d_sums && d_counts, "null pointer");
	CS_REQUIRE(feat_norm2_max >= 0.0 && feat_norm2_max < 1e15, "feat_norm2_max out of range");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(aligned16(d_f0) && aligned16(d_f1) && aligned16(d_f2), "feature planes must be 16-byte aligned");
	CS_REQUIRE(!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 3u) == 0, "labels must be 4-byte aligned");
	CS_REQUIRE(!d_centers_out || d_stats, "fused finalize needs d_stats");
	CS_REQUIRE(d_centers_out != d_centers_in, "centers_in and centers_out must not alias");
	LloydParams p{};
	p.f0 = d_f0; p.f1 = d_f1; p.f2 = d_f2; p.n = n; p.centers = d_centers_in; p.K = K;
	p.x2max = feat_norm2_max;
	p.labels = d_labels; p.sums = d_sums; p.counts = d_counts; p.inertia = nullptr;
	p.partials = ctx->d_partials; p.counter = ctx->d_counter;
	p.centers_out = d_centers_out; p.stats = d_stats;
	return launch_k<FM_F32>(ctx, p, flags, (cudaStream_t)stream);
}

extern "C" int cs_lloyd_step_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1,
                                 const float *d_f2, int64_t n, const double *d_centers, int K,
                                 uint8_t *d_labels, double *d_sums, double *d_counts,
                                 double *d_inertia, double feat_norm2_max, int flags, void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_centers && d_sums && d_counts, "null pointer");
	CS_REQUIRE(feat_norm2_max >= 0.0 && feat_norm2_max < 1e15, "feat_norm2_max out of range");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(aligned16(d_f0) && aligned16(d_f1) && aligned16(d_f2), "feature planes must be 16-byte aligned");
	CS_REQUIRE(!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 3u) == 0, "labels must be 4-byte aligned");
	LloydParams p{};
	p.f0 = d_f0; p.f1 = d_f1; p.f2 = d_f2; p.n = n; p.centers = d_centers; p.K = K;
	p.x2max = feat_norm2_max;
	p.labels = d_labels; p.sums = d_sums; p.counts = d_counts; p.inertia = d_inertia;
	p.partials = ctx->d_partials; p.counter = ctx->d_counter;
	return launch_k<FM_F32>(ctx, p, flags, (cudaStream_t)stream);
}

static int lloyd_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3, int mask_mode,
                     int min_bright, double x2max, const double *d_centers, int K, uint8_t *d_labels,
                     double *d_sums, double *d_counts, double *d_inertia, double *d_centers_out,
                     double *d_stats, int flags, void *stream) {
	CS_REQUIRE(ctx && d_px && d_centers && d_sums && d_counts, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(aligned16(d_px), "pixels must be 16-byte aligned");
	CS_REQUIRE(!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 3u) == 0, "labels must be 4-byte aligned");
	CS_REQUIRE(!d_centers_out || (d_stats && d_centers_out != d_centers), "fused finalize needs d_stats and distinct centre buffers");
	if (!d_lut3) {
		// exact-integer contract of the RGB path (include/colorsimplify.h): the lane-private fp32 slots must stay below 2^24
		// even if every pixel falls into one cluster (x4 for an uneven spread over the slots)
		int kp = 8;
		while (kp < K) kp <<= 1;
		const int copies = (kAccBytesPerWarp / (kp * 16)) > 32 ? 32 : (kAccBytesPerWarp / (kp * 16));
		const long long tile = kp <= 16 ? VarSmallK::TILE : VarLargeK::TILE;
		const long long ntiles = (n + tile - 1) / tile;
		const long long ctas = ctx->launch_images > 1 ? ctx->launch_ctas_per_image : (ntiles < ctx->sm_count ? (ntiles < 1 ? 1 : ntiles) : ctx->sm_count);
		CS_REQUIRE((double)n * 255.0 * 4.0 <= 16777216.0 * (double)(ctas * 16 * copies),
		           "image too large for exact fp32 slot sums at this K (see the exactness limit in colorsimplify.h)");
	}
	LloydParams p{};
	p.rgba = reinterpret_cast<const uint32_t *>(d_px); p.n = n; p.min_rgb_sum = min_bright;
	p.lut3 = d_lut3; p.mask_mode = mask_mode; p.x2max = x2max;
	p.centers = d_centers; p.K = K; p.labels = d_labels; p.sums = d_sums; p.counts = d_counts;
	p.inertia = d_inertia; p.partials = ctx->d_partials; p.counter = ctx->d_counter;
	p.centers_out = d_centers_out; p.stats = d_stats;
	return launch_k<FM_RGBA8>(ctx, p, flags, (cudaStream_t)stream);
}

extern "C" int cs_lloyd_step_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                                   const double *d_centers, int K, uint8_t *d_labels,
                                   double *d_sums, double *d_counts, double *d_inertia, int flags,
                                   void *stream) {
	return lloyd_px8(ctx, d_rgba, n, nullptr, 0, min_rgb_sum, 3.0 * 255.0 * 255.0, d_centers, K, d_labels, d_sums,
	                 d_counts, d_inertia, nullptr, nullptr, flags, stream);
}

extern "C" int cs_lloyd_iter_rgba8(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n, int min_rgb_sum,
                                   const double *d_centers_in, int K, uint8_t *d_labels, double *d_sums,
                                   double *d_counts, double *d_centers_out, double *d_stats, int flags,
                                   void *stream) {
	CS_REQUIRE(d_centers_out && d_stats, "null pointer");
	return lloyd_px8(ctx, d_rgba, n, nullptr, 0, min_rgb_sum, 3.0 * 255.0 * 255.0, d_centers_in, K, d_labels, d_sums,
	                 d_counts, nullptr, d_centers_out, d_stats, flags, stream);
}

extern "C" int cs_lloyd_step_px8lut(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3,
                                    int mask_mode, int min_bright, double feat_norm2_max,
                                    const double *d_centers, int K, uint8_t *d_labels, double *d_sums,
                                    double *d_counts, double *d_inertia, double *d_centers_out,
                                    double *d_stats, int flags, void *stream) {
	CS_REQUIRE(d_lut3, "null pointer");
	CS_REQUIRE(mask_mode == 0 || mask_mode == 1, "mask_mode must be 0 or 1");
	CS_REQUIRE(feat_norm2_max >= 0.0 && feat_norm2_max < 1e15, "feat_norm2_max out of range");
	return lloyd_px8(ctx, d_px, n, d_lut3, mask_mode, min_bright, feat_norm2_max, d_centers, K, d_labels, d_sums,
	                 d_counts, d_inertia, d_centers_out, d_stats, flags, stream);
}

extern "C" int cs_lloyd_set_feature_box(cs_ctx *ctx, const double *h_lo3, const double *h_hi3) {
	CS_REQUIRE(ctx, "null pointer");
	if (!h_lo3 || !h_hi3) { ctx->box_set = 0; return 0; }
	for (int j = 0; j < 3; ++j) {
		CS_REQUIRE(h_lo3[j] <= h_hi3[j] && h_hi3[j] - h_lo3[j] < 1e15, "box must have lo <= hi (finite)");
		ctx->box_lo[j] = h_lo3[j]; ctx->box_hi[j] = h_hi3[j];
	}
	ctx->box_set = 1;
	return 0;
}

extern "C" int cs_lloyd_set_grid_policy(cs_ctx *ctx, int policy) {
	CS_REQUIRE(ctx && policy >= -1 && policy <= 1, "policy must be -1, 0 or 1");
	ctx->grid_policy = policy;
	return 0;
}

// ================= whole KMeans fits of a SAMPLE in one launch (fp64 rows resident in shared memory) =================
// simplify_colors_perceptual_fast fits its palette with KMeans(n_clusters=K, random_state=42, n_init=10,
// max_iter=100) on the <= 5000 distinct colours of a sample (app/processing/color_simplify.py:669-675).  One CTA per
// initialisation runs the complete loop of _kmeans_single_lloyd (sklearn/cluster/_kmeans.py:705-758) on the rows,
// which it keeps in shared memory (24 B per row): E-step (fp64 direct distances, strict < = first minimum), per-cluster
// sums in a fixed order (one warp per cluster, lane-strided, butterfly), relocation of empty clusters
// (_k_means_common.pyx:167-211: farthest points first, lowest index on equal distances), the M-step tail shared with
// the per-pixel kernels (finalize_block), the stop rule (labels repeat -> strict convergence; else
// sum(shift^2) <= tol), then the final E-step and the inertia.  No host round trip inside a fit; the n_init fits
// run side by side.  Run-to-run deterministic.
constexpr int kSmallFitMaxRows = 5120;
constexpr int kSmallFitThreads = 1024;

__device__ __forceinline__ double d2_rows(const double *c, double x, double y, double z) {
	const double dx = x - c[0], dy = y - c[1], dz = z - c[2];
	return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(kSmallFitThreads) fit_rows64_small_kernel(
    const double *__restrict__ g_rows, int n, const double *__restrict__ g_inits, int K, int max_iter, double tol,
    double *__restrict__ g_centers, uint8_t *__restrict__ g_labels, double *__restrict__ g_stats) {
	extern __shared__ __align__(16) unsigned char fs_smem[];
	double *rows = reinterpret_cast<double *>(fs_smem);       // n x 3
	double *dist = rows + 3 * (size_t)kSmallFitMaxRows;       // n
	double *cen = dist + kSmallFitMaxRows;                    // 2 x K x 3 (current / next)
	double *sums = cen + 2 * CS_MAX_K * 3;                    // K x 3
	double *counts = sums + CS_MAX_K * 3;                     // K
	double *shift2 = counts + CS_MAX_K;                       // K
	double *red = shift2 + CS_MAX_K;                          // 32
	uint8_t *lab = reinterpret_cast<uint8_t *>(red + 32);     // 2 x n (this iteration / the one before)
	__shared__ double s_shift2, s_stats[4];
	__shared__ int s_nempty, s_stop;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5, b = blockIdx.x;
	for (int i = t; i < 3 * n; i += kSmallFitThreads) rows[i] = g_rows[i];
	for (int i = t; i < 3 * K; i += kSmallFitThreads) cen[i] = g_inits[(size_t)b * K * 3 + i];
	__syncthreads();
	int cur = 0, lcur = 0, n_iter = 0;
	auto estep = [&](const double *c, uint8_t *l) {
		for (int i = t; i < n; i += kSmallFitThreads) {
			const double x = rows[3 * i], y = rows[3 * i + 1], z = rows[3 * i + 2];
			double best = 1e300;
			int bk = 0;
			for (int k = 0; k < K; ++k) {
				const double d = d2_rows(c + 3 * k, x, y, z);
				if (d < best) { best = d; bk = k; }
			}
			l[i] = (uint8_t)bk;
			dist[i] = best;
		}
	};
	for (int it = 1; it <= max_iter; ++it) {
		const double *c = cen + cur * CS_MAX_K * 3;
		double *cn = cen + (cur ^ 1) * CS_MAX_K * 3;
		uint8_t *l = lab + (size_t)lcur * kSmallFitMaxRows, *lo = lab + (size_t)(lcur ^ 1) * kSmallFitMaxRows;
		estep(c, l);
		__syncthreads();
		// per-cluster sums, fixed order: warp w takes clusters w, w + 32, ...
		for (int k = warp; k < K; k += kSmallFitThreads / 32) {
			double s0 = 0.0, s1 = 0.0, s2 = 0.0, cw = 0.0;
			for (int i = lane; i < n; i += 32)
				if (l[i] == k) { s0 += rows[3 * i]; s1 += rows[3 * i + 1]; s2 += rows[3 * i + 2]; cw += 1.0; }
			for (int o = 16; o > 0; o >>= 1) {
				s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
				s2 += __shfl_xor_sync(0xffffffffu, s2, o); cw += __shfl_xor_sync(0xffffffffu, cw, o);
			}
			if (lane == 0) { sums[3 * k] = s0; sums[3 * k + 1] = s1; sums[3 * k + 2] = s2; counts[k] = cw; }
		}
		__syncthreads();
		if (t == 0) {
			int ne = 0;
			for (int k = 0; k < K; ++k) ne += counts[k] == 0.0;
			s_nempty = ne;
			if (ne) {
				// _relocate_empty_clusters_dense: the farthest points (from their own centres) move into the empty
				// clusters, in index order of the clusters; nothing moves when no point is away from its centre
				double dmax = 0.0, xxmax = 1.0;
				for (int i = 0; i < n; ++i) {
					dmax = fmax(dmax, dist[i]);
					xxmax = fmax(xxmax, rows[3 * i] * rows[3 * i] + rows[3 * i + 1] * rows[3 * i + 1] + rows[3 * i + 2] * rows[3 * i + 2]);
				}
				if (dmax > 1e-24 * xxmax) {
					for (int e = 0; e < K; ++e) {
						if (counts[e] != 0.0) continue;
						int far = -1;
						double fd = -1.0;
						for (int i = 0; i < n; ++i)
							if (dist[i] > fd) { fd = dist[i]; far = i; }
						if (far < 0) break;
						dist[far] = -2.0;  // taken
						const int old = l[far];
						for (int j = 0; j < 3; ++j) { sums[3 * old + j] -= rows[3 * far + j]; sums[3 * e + j] = rows[3 * far + j]; }
						counts[e] = 1.0;
						counts[old] -= 1.0;
					}
				}
			}
		}
		__syncthreads();
		double sh2 = 0.0;
		int nempty = 0;
		finalize_block(sums, counts, c, K, cn, s_stats, shift2, sh2, nempty);
		if (t == 0) s_shift2 = sh2;
		// labels equal to those of the iteration before?  (never on the first iteration)
		int same = it > 1;
		for (int i = t; i < n && same; i += kSmallFitThreads) same = l[i] == lo[i];
		same = __syncthreads_and(same);  // (also publishes s_shift2 and the new centres)
		cur ^= 1;
		n_iter = it;
		if (same || s_shift2 <= tol) {
			if (t == 0) s_stop = same ? 2 : 1;
			break;
		}
		lcur ^= 1;
		if (t == 0) s_stop = 0;
	}
	__syncthreads();
	// labels and inertia for the final centres (_kmeans_single_lloyd :740-758)
	const double *c = cen + cur * CS_MAX_K * 3;
	uint8_t *l = lab + (size_t)lcur * kSmallFitMaxRows;
	estep(c, l);
	__syncthreads();
	double acc = 0.0;
	for (int i = t; i < n; i += kSmallFitThreads) acc += dist[i];
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) red[warp] = acc;
	__syncthreads();
	if (t == 0) {
		double s = 0.0;
		for (int w = 0; w < kSmallFitThreads / 32; ++w) s += red[w];
		g_stats[4 * b] = s;
		g_stats[4 * b + 1] = (double)n_iter;
		g_stats[4 * b + 2] = (double)s_stop;
		g_stats[4 * b + 3] = (double)s_nempty;
	}
	for (int i = t; i < 3 * K; i += kSmallFitThreads) g_centers[(size_t)b * K * 3 + i] = c[i];
	for (int i = t; i < n; i += kSmallFitThreads) g_labels[(size_t)b * n + i] = l[i];
}

extern "C" int cs_kmeans_fit_rows64_small(cs_ctx *ctx, const double *d_rows, int64_t n, const double *d_inits, int n_init, int K,
                                          int max_iter, double tol, double *d_centers, uint8_t *d_labels, double *d_stats,
                                          void *stream) {
	CS_REQUIRE(ctx && d_rows && d_inits && d_centers && d_labels && d_stats, "null pointer");
	CS_REQUIRE(n >= 1 && n <= kSmallFitMaxRows, "n must be in [1, 5120] (rows are kept in shared memory)");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n_init >= 1 && n_init <= 1024 && max_iter >= 1, "bad K, n_init or max_iter");
	const size_t smem = sizeof(double) * (4 * (size_t)kSmallFitMaxRows + 2 * CS_MAX_K * 3 + CS_MAX_K * 3 + 2 * CS_MAX_K + 32) +
	                    2 * (size_t)kSmallFitMaxRows;
	static bool attr_set = false;
	if (!attr_set) {
		CS_CUDA(cudaFuncSetAttribute(fit_rows64_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		attr_set = true;
	}
	fit_rows64_small_kernel<<<n_init, kSmallFitThreads, smem, (cudaStream_t)stream>>>(d_rows, (int)n, d_inits, K, max_iter, tol,
	                                                                               d_centers, d_labels, d_stats);
	CS_CUDA(cudaGetLastError());
	return 0;
}

extern "C" int cs_lloyd_finalize(cs_ctx *ctx, const double *d_sums, const double *d_counts,
                                 const double *d_centers_old, int K, double *d_centers_new,
                                 double *d_stats, void *stream) {
	CS_REQUIRE(ctx && d_sums && d_counts && d_centers_old && d_centers_new && d_stats, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_sums, d_counts, d_centers_old, K,
	                                                     d_centers_new, d_stats);
	CS_CUDA(cudaGetLastError());
	return 0;
}

// Multi-GPU fused iteration: as cs_lloyd_iter_f32, but the last CTA exchanges this rank's partial with
// every peer through the cudaIpc-mapped mailboxes (mg.cu) and reduces all ranks' partials in rank order
// before the M-step tail — assign + update + "all-reduce" + finalize in ONE kernel, no collective launch.
extern "C" int cs_lloyd_iter_f32_mg(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2,
                                    int64_t n, const double *d_centers_in, int K, uint8_t *d_labels,
                                    double *d_sums, double *d_counts, double *d_centers_out, double *d_stats,
                                    double feat_norm2_max, int flags, void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_centers_in && d_sums && d_counts && d_centers_out && d_stats, "null pointer");
	CS_REQUIRE(ctx->mg_own && ctx->mg_world > 1, "no connected mailbox: call cs_mg_create / cs_mg_connect first");
	CS_REQUIRE(feat_norm2_max >= 0.0 && feat_norm2_max < 1e15, "feat_norm2_max out of range");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n >= 0, "n must be >= 0");
	CS_REQUIRE(aligned16(d_f0) && aligned16(d_f1) && aligned16(d_f2), "feature planes must be 16-byte aligned");
	CS_REQUIRE(!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 3u) == 0, "labels must be 4-byte aligned");
	CS_REQUIRE(d_centers_out != d_centers_in, "centers_in and centers_out must not alias");
	LloydParams p{};
	p.f0 = d_f0; p.f1 = d_f1; p.f2 = d_f2; p.n = n; p.centers = d_centers_in; p.K = K;
	p.x2max = feat_norm2_max;
	p.labels = d_labels; p.sums = d_sums; p.counts = d_counts; p.inertia = nullptr;
	p.partials = ctx->d_partials; p.counter = ctx->d_counter;
	p.centers_out = d_centers_out; p.stats = d_stats;
	for (int q = 0; q < ctx->mg_world; ++q) p.mb[q] = ctx->mg_peer[q];
	p.world = ctx->mg_world; p.rank = ctx->mg_rank; p.epoch = ++ctx->mg_epoch;
	return launch_k<FM_F32>(ctx, p, flags, (cudaStream_t)stream);
}

// Batched fused iteration on packed RGBA8 images: `n_images` independent k-means problems (one K x 3
// table, one label map, one set of accumulators per image) in ONE launch, gridDim.y = image —
// BASELINE config 4 (a batch of 1920x1080 images, images partitioned across GPUs, no collective).
extern "C" int cs_lloyd_iter_rgba8_batched(cs_ctx *ctx, const uint8_t *d_rgba, int64_t n_per_image, int n_images,
                                           int min_rgb_sum, const double *d_centers_in, int K, uint8_t *d_labels,
                                           double *d_sums, double *d_counts, double *d_centers_out, double *d_stats,
                                           int flags, void *stream) {
	CS_REQUIRE(ctx && d_rgba && d_centers_in && d_sums && d_counts && d_centers_out && d_stats, "null pointer");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K, "K must be in [1,256]");
	CS_REQUIRE(n_per_image > 0 && (n_per_image & 3) == 0, "n_per_image must be a positive multiple of 4");
	CS_REQUIRE(n_images >= 1, "n_images must be >= 1");
	CS_REQUIRE(aligned16(d_rgba), "pixels must be 16-byte aligned");
	CS_REQUIRE(!d_labels || (reinterpret_cast<uintptr_t>(d_labels) & 3u) == 0, "labels must be 4-byte aligned");
	CS_REQUIRE(d_centers_out != d_centers_in, "centers_in and centers_out must not alias");
	// CTAs per image: enough images x CTAs for >= 4 waves over the SMs, bounded by the partial scratch
	int per = (4 * ctx->sm_count + n_images - 1) / n_images;
	if (per < 1) per = 1;
	if (per > 16) per = 16;
	const int chunk_max = kMaxPartialBlocks / per < kMaxBatchImages ? kMaxPartialBlocks / per : kMaxBatchImages;
	for (int i0 = 0; i0 < n_images; i0 += chunk_max) {
		const int cnt = n_images - i0 < chunk_max ? n_images - i0 : chunk_max;
		LloydParams p{};
		p.rgba = reinterpret_cast<const uint32_t *>(d_rgba) + (long long)i0 * n_per_image;
		p.n = n_per_image; p.min_rgb_sum = min_rgb_sum; p.lut3 = nullptr; p.mask_mode = 0;
		p.x2max = 3.0 * 255.0 * 255.0;
		p.centers = d_centers_in + (size_t)i0 * K * 3; p.K = K;
		p.labels = d_labels ? d_labels + (long long)i0 * n_per_image : nullptr;
		p.sums = d_sums + (size_t)i0 * K * 3; p.counts = d_counts + (size_t)i0 * K; p.inertia = nullptr;
		p.partials = ctx->d_partials; p.counter = ctx->d_counter;
		p.centers_out = d_centers_out + (size_t)i0 * K * 3; p.stats = d_stats + (size_t)i0 * 4;
		p.img_stride_px = n_per_image; p.img_stride_label = n_per_image;
		ctx->launch_images = cnt; ctx->launch_ctas_per_image = per;
		const int rc = launch_k<FM_RGBA8>(ctx, p, flags, (cudaStream_t)stream);
		ctx->launch_images = 1;
		if (rc) return rc;
	}
	return 0;
}

// Queue `n_launch` fused iterations back to back, ping-ponging the centre buffers (a -> b, b -> a, ...),
// with device-side loop control: see LloydParams::ctl.  Replaces the per-iteration host round trip of
// _kmeans_single_lloyd's loop (sklearn/cluster/_kmeans.py:705-738) by one check per batch.
extern "C" int cs_lloyd_run_f32(cs_ctx *ctx, const float *d_f0, const float *d_f1, const float *d_f2, int64_t n,
                                double *d_centers_a, double *d_centers_b, int K, double *d_sums, double *d_counts,
                                double *d_stats, double feat_norm2_max, int flags, int n_launch, double *d_ctl,
                                void *stream) {
	CS_REQUIRE(ctx && d_f0 && d_f1 && d_f2 && d_centers_a && d_centers_b && d_sums && d_counts && d_stats && d_ctl, "null pointer");
	CS_REQUIRE(d_centers_a != d_centers_b, "the two centre buffers must differ");
	CS_REQUIRE(feat_norm2_max >= 0.0 && feat_norm2_max < 1e15, "feat_norm2_max out of range");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0 && n_launch >= 0, "bad K, n or n_launch");
	CS_REQUIRE(aligned16(d_f0) && aligned16(d_f1) && aligned16(d_f2), "feature planes must be 16-byte aligned");
	for (int i = 0; i < n_launch; ++i) {
		LloydParams p{};
		p.f0 = d_f0; p.f1 = d_f1; p.f2 = d_f2; p.n = n; p.K = K;
		p.centers = (i & 1) ? d_centers_b : d_centers_a;
		p.centers_out = (i & 1) ? d_centers_a : d_centers_b;
		p.x2max = feat_norm2_max;
		p.labels = nullptr; p.sums = d_sums; p.counts = d_counts; p.inertia = nullptr;
		p.partials = ctx->d_partials; p.counter = ctx->d_counter; p.stats = d_stats; p.ctl = d_ctl;
#ifdef CS_PHASE_TIMING
		p.phase_ts = ctx->d_scratch64 + 16;
#endif
		const int rc = launch_k<FM_F32>(ctx, p, i > 0 ? (flags | CS_LLOYD_CHAINED) : flags, (cudaStream_t)stream);
		if (rc) return rc;
	}
	return 0;
}

// same on packed 4 x u8 pixels (d_lut3 NULL = RGB features; mask as cs_lloyd_step_px8lut)
extern "C" int cs_lloyd_run_px8(cs_ctx *ctx, const uint8_t *d_px, int64_t n, const float *d_lut3, int mask_mode,
                                int min_bright, double feat_norm2_max, double *d_centers_a, double *d_centers_b, int K,
                                double *d_sums, double *d_counts, double *d_stats, int flags, int n_launch,
                                double *d_ctl, void *stream) {
	CS_REQUIRE(ctx && d_px && d_centers_a && d_centers_b && d_sums && d_counts && d_stats && d_ctl, "null pointer");
	CS_REQUIRE(d_centers_a != d_centers_b, "the two centre buffers must differ");
	CS_REQUIRE(mask_mode == 0 || mask_mode == 1, "mask_mode must be 0 or 1");
	CS_REQUIRE(K >= 1 && K <= CS_MAX_K && n >= 0 && n_launch >= 0, "bad K, n or n_launch");
	CS_REQUIRE(aligned16(d_px), "pixels must be 16-byte aligned");
	for (int i = 0; i < n_launch; ++i) {
		LloydParams p{};
		p.rgba = reinterpret_cast<const uint32_t *>(d_px); p.n = n; p.min_rgb_sum = min_bright;
		p.lut3 = d_lut3; p.mask_mode = mask_mode; p.x2max = d_lut3 ? feat_norm2_max : 3.0 * 255.0 * 255.0;
		p.K = K;
		p.centers = (i & 1) ? d_centers_b : d_centers_a;
		p.centers_out = (i & 1) ? d_centers_a : d_centers_b;
		p.labels = nullptr; p.sums = d_sums; p.counts = d_counts; p.inertia = nullptr;
		p.partials = ctx->d_partials; p.counter = ctx->d_counter; p.stats = d_stats; p.ctl = d_ctl;
		const int rc = launch_k<FM_RGBA8>(ctx, p, i > 0 ? (flags | CS_LLOYD_CHAINED) : flags, (cudaStream_t)stream);
		if (rc) return rc;
	}
	return 0;
}
