// cs_common.cuh — shared internals of libcolorsimplify (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/colorsimplify.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcolorsimplify is written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace cs {

void set_error(const char *fmt, ...);

#define CS_CUDA(expr)                                                                       \
	do {                                                                                    \
		cudaError_t _e = (expr);                                                            \
		if (_e != cudaSuccess) {                                                            \
			cs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
			              cudaGetErrorString(_e));                                          \
			return (int)_e;                                                                 \
		}                                                                                   \
	} while (0)

#define CS_REQUIRE(cond, msg)                                                               \
	do {                                                                                    \
		if (!(cond)) {                                                                      \
			cs::set_error("%s: %s", __func__, msg);                                         \
			return (int)CS_ERR_ARG;                                                         \
		}                                                                                   \
	} while (0)

constexpr int kMaxPartialBlocks = 1024;  // upper bound on persistent grid size
constexpr int kMaxPartialVals = CS_MAX_K * 4 + 8;
constexpr int kMaxBatchImages = 1024;    // images per batched launch (one "blocks finished" counter each)
// cell grid of the grid-filtered assignment (lloyd.cu): table capacity in cells (one u32 of four candidate
// bytes per cell) and the pool of 8-candidate entries for the cells that need more than four
constexpr int kGridCap = 9984;
constexpr int kGridPool = 768;  // (K <= 16 uses the first 256: the pool index must fit two label bytes)
constexpr int kGridWords = kGridCap + 2 * kGridPool;  // u32 words bulk-copied into shared memory
// behind the table in d_grid: four pool counters (two tiers, ping-pong by build epoch), then one bit per cell set by
// the Lloyd kernel when a pixel of the cell had to take the all-K walk (a crowded cell that got no pool entry): the
// next build serves those cells first
constexpr int kGridTier1 = 512;                        // pool entries reserved for marked cells
constexpr int kGridMarkWords = (kGridCap + 31) / 32;
constexpr int kGridAllWords = kGridWords + 4 + kGridMarkWords;

} // namespace cs

// Multi-GPU mailbox (lloyd.cu, mg.cu): every rank owns one in its own HBM and maps its peers' over
// cudaIpc (NVLink P2P).  Flag-in-data protocol (as NCCL's LL): a double travels as two 8-byte words
// {32 data bits, 32-bit epoch tag}, each written with ONE scalar 8-byte store, so a word is either old or
// complete and carries its own "valid for epoch e" mark.  Writer r stores its per-iteration partial into
// slot [parity][r] of EVERY rank's mailbox; a reader polls the words of its own mailbox until their tags
// show the epoch.  No separate flag and no system-scope fence: a release at .sys scope also waits for the
// rank's own 64 MB of freshly written labels to drain, which cost 25-40 us per iteration at 8 GPUs.
namespace cs { constexpr int kMgMaxRanks = 8; }
struct cs_mailbox {
	unsigned long long word[2][cs::kMgMaxRanks][2 * cs::kMaxPartialVals];  // (epoch tag << 32) | data half
	unsigned long long error;  // set to the epoch of a wait that timed out
};

struct cs_ctx {
	int device;
	int sm_count;
	// multi-GPU exchange state (null / 1 when single-GPU)
	cs_mailbox *mg_own;
	cs_mailbox *mg_peer[cs::kMgMaxRanks];  // [rank] = own, others = cudaIpc-mapped
	int mg_world, mg_rank;
	unsigned long long mg_epoch;
	double *d_partials;        // [kMaxPartialBlocks][kMaxPartialVals] scratch shared by the scan / histogram kernels
	// per-CTA partials of the Lloyd kernel as self-validating words ((launch epoch << 32) | half of a double):
	// [kMaxPartialBlocks][2 * kMaxPartialVals], written by nothing else, zeroed at creation (epochs start at 1)
	unsigned long long *d_partial_words;
	mutable unsigned long long lloyd_epoch;
	// grid-filtered assignment: candidate table + overflow pool (kGridWords u32) followed by two pool
	// counters (ping-pong by build epoch); the feature box the caller vouched for (cs_lloyd_set_feature_box)
	uint32_t *d_grid;
	unsigned long long grid_epoch;
	double box_lo[3], box_hi[3];
	int box_set;
	int grid_policy;  // 0 = where it is faster (K > 16), 1 = wherever eligible (4 <= K <= 64), -1 = never
	uint32_t *d_remap_tab;     // K4 grid path: 32^3 candidate entries over the RGB cube (lazily allocated)
	uint8_t *d_remap_lut;      // K4 colour-table path: 2^24 label bytes (rows of the mixed cells only) + the 128 KB fine-cell table (lazily allocated)
	int remap_policy;          // cs_remap_set_policy: 0 = by image size, 1 = three-phase tiles, 2 = colour table, -1 = direct kernel
	unsigned int *d_counter;   // "blocks finished" counters for the last-block combine (one per image of a batched launch)
	int launch_images, launch_ctas_per_image;  // set around a batched launch (1 otherwise)
	unsigned long long *d_scratch64; // 64 u64 of misc scratch (relocation keys, ...)
	// persistent device buffers for the host-buffer convenience path
	void *d_host_buf;
	size_t host_buf_bytes;
	// host-buffer path: uploads run on host_copy, kernels on host_comp (both non-blocking, created on first
	// use); chunk i of the upload signals host_ev[i] so that its LAB conversion overlaps the next chunk's copy
	cudaStream_t host_copy, host_comp;
	cudaEvent_t host_ev[8];
	void *host_stager;  // cs_host_upload: page-locked staging ring + host copy threads (hostio.cu, created on first use)
};

namespace cs {
void host_stager_destroy(cs_ctx *ctx);  // hostio.cu
}

namespace cs {

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA without a tensor map) -----------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
	             "r"(bytes)
	             : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "CS_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra CS_DONE;\n"
	    "bra CS_WAIT;\n"
	    "CS_DONE:\n"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// non-blocking look at a barrier phase (used one tile ahead, so that the result's latency hides behind
// the current tile's arithmetic); a true result orders the bulk-copy data before later loads like a wait
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
	uint32_t done;
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
	    "selp.u32 %0, 1, 0, P1;\n"
	    "}"
	    : "=r"(done)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return done != 0;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
// same wait with a sleep between polls: for the single producer lane, whose spin would otherwise
// take issue slots from the four consumer warps of its scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
	uint32_t done = 0;
	while (true) {
		asm volatile(
		    "{\n"
		    ".reg .pred P1;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, P1;\n"
		    "}"
		    : "=r"(done)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (done) break;
		__nanosleep(100);
	}
}
// global -> shared bulk copy, completion signalled on `bar` (complete_tx of `bytes`).
// dst, src 16-byte aligned, bytes a positive multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                         uint64_t *bar) {
	asm volatile(
	    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
	        "r"(smem_u32(dst)),
	    "l"(src), "r"(bytes), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream_u4(uint4 *p, const uint4 &v) {
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
	             "r"(v.y), "r"(v.z), "r"(v.w)
	             : "memory");
}

// Pixel selection shared by the masked kernels.  mask_mode 0: alpha > 0 && b0+b1+b2 > min_bright
// (color_simplify.py:44, 56-64: mean(rgb) > 30 <=> r+g+b > 90); 1: alpha > 0 && b2 > min_bright
// (the V filter of :956-963 on HSVA pixels).  min_bright < 0 keeps every opaque pixel.
__device__ __forceinline__ bool px_selected(uint32_t w, int mask_mode, int min_bright) {
	if (!(w >> 24)) return false;
	const int br = mask_mode == 0 ? (int)((w & 0xFFu) + ((w >> 8) & 0xFFu) + ((w >> 16) & 0xFFu))
	                              : (int)((w >> 16) & 0xFFu);
	return br > min_bright;
}
// validity of a labelled pixel: by the selection pixels when given, else by the 255 sentinel
__device__ __forceinline__ bool label_valid(const uint32_t *selpx, long long i, int mask_mode, int min_bright,
                                            uint32_t label, int K) {
	return selpx ? px_selected(selpx[i], mask_mode, min_bright) : (label < (uint32_t)K && (K == 256 || label != 255u));
}

inline int grid_for(const cs_ctx *ctx, int64_t work_items, int per_sm) {
	int64_t g = (int64_t)ctx->sm_count * per_sm;
	if (work_items < g) g = work_items < 1 ? 1 : work_items;
	return (int)g;
}

} // namespace cs
