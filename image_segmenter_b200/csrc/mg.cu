// mg.cu — multi-GPU exchange plumbing: per-rank mailbox in device memory, exported / mapped with
// cudaIpc so that the Lloyd kernel of every rank can store its per-iteration partial straight into
// its peers' HBM over NVLink (lloyd.cu, cs_lloyd_iter_f32_mg).  One process per GPU; the 64-byte
// handles travel between the processes by whatever the host has (torch.distributed all_gather).
#include <string.h>

#include "cs_common.cuh"

using namespace cs;

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");

extern "C" int cs_mg_create(cs_ctx *ctx, int world, int rank, void *h_handle64) {
	CS_REQUIRE(ctx && h_handle64, "null pointer");
	CS_REQUIRE(world >= 1 && world <= kMgMaxRanks && rank >= 0 && rank < world, "world must be in [1,8], rank in [0,world)");
	CS_REQUIRE(!ctx->mg_own, "mailbox already created");
	CS_CUDA(cudaSetDevice(ctx->device));
	CS_CUDA(cudaMalloc(&ctx->mg_own, sizeof(cs_mailbox)));
	CS_CUDA(cudaMemset(ctx->mg_own, 0, sizeof(cs_mailbox)));
	CS_CUDA(cudaDeviceSynchronize());
	cudaIpcMemHandle_t h;
	CS_CUDA(cudaIpcGetMemHandle(&h, ctx->mg_own));
	memcpy(h_handle64, &h, sizeof(h));
	ctx->mg_world = world; ctx->mg_rank = rank; ctx->mg_epoch = 0;
	for (int q = 0; q < kMgMaxRanks; ++q) ctx->mg_peer[q] = nullptr;
	ctx->mg_peer[rank] = ctx->mg_own;
	return 0;
}

extern "C" int cs_mg_connect(cs_ctx *ctx, const void *h_handles) {
	CS_REQUIRE(ctx && h_handles && ctx->mg_own, "null pointer or no mailbox");
	CS_CUDA(cudaSetDevice(ctx->device));
	for (int q = 0; q < ctx->mg_world; ++q) {
		if (q == ctx->mg_rank) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, static_cast<const unsigned char *>(h_handles) + 64 * q, sizeof(h));
		void *p = nullptr;
		CS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		ctx->mg_peer[q] = static_cast<cs_mailbox *>(p);
	}
	return 0;
}

// epoch of a timed-out wait (0 = none); synchronises the device
extern "C" int cs_mg_error(cs_ctx *ctx, unsigned long long *h_epoch) {
	CS_REQUIRE(ctx && h_epoch && ctx->mg_own, "null pointer or no mailbox");
	CS_CUDA(cudaMemcpy(h_epoch, &ctx->mg_own->error, sizeof(*h_epoch), cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int cs_mg_destroy(cs_ctx *ctx) {
	if (!ctx || !ctx->mg_own) return 0;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (int q = 0; q < ctx->mg_world; ++q)
		if (q != ctx->mg_rank && ctx->mg_peer[q]) cudaIpcCloseMemHandle(ctx->mg_peer[q]);
	cudaFree(ctx->mg_own);
	ctx->mg_own = nullptr;
	ctx->mg_world = 1;
	return 0;
}
