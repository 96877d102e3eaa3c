// api.cu — context, error reporting and the host-buffer convenience path of libcolorsimplify.
#include <stdarg.h>
#include <string.h>

#include "cs_common.cuh"

namespace cs {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
} // namespace cs

using namespace cs;

extern "C" int cs_abi_version(void) { return CS_ABI_VERSION; }

extern "C" const char *cs_last_error(void) { return g_err; }

extern "C" int cs_ctx_create(int device, cs_ctx **out) {
	if (!out) {
		set_error("cs_ctx_create: out is null");
		return CS_ERR_ARG;
	}
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		set_error("cs_ctx_create: no CUDA device (%s) — libcolorsimplify has no CPU path",
		          cudaGetErrorString(e));
		return CS_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= count) {
		set_error("cs_ctx_create: device %d out of range [0,%d)", device, count);
		return CS_ERR_ARG;
	}
	CS_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CS_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		set_error("cs_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
		          device, prop.major, prop.minor);
		return CS_ERR_NO_DEVICE;
	}
	cs_ctx *c = new cs_ctx();
	memset(c, 0, sizeof(*c));
	c->device = device;
	c->mg_world = 1;
	c->launch_images = 1;
	c->sm_count = prop.multiProcessorCount;
	CS_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * (size_t)kMaxPartialBlocks * kMaxPartialVals));
	CS_CUDA(cudaMalloc(&c->d_partial_words, sizeof(unsigned long long) * (size_t)kMaxPartialBlocks * 2 * kMaxPartialVals));
	CS_CUDA(cudaMemset(c->d_partial_words, 0, sizeof(unsigned long long) * (size_t)kMaxPartialBlocks * 2 * kMaxPartialVals));
	CS_CUDA(cudaMalloc(&c->d_grid, sizeof(uint32_t) * kGridAllWords));
	CS_CUDA(cudaMemset(c->d_grid, 0, sizeof(uint32_t) * kGridAllWords));
	CS_CUDA(cudaMalloc(&c->d_counter, sizeof(unsigned int) * kMaxBatchImages));
	CS_CUDA(cudaMemset(c->d_counter, 0, sizeof(unsigned int) * kMaxBatchImages));
	CS_CUDA(cudaMalloc(&c->d_scratch64, 64 * sizeof(unsigned long long)));
	CS_CUDA(cudaMemset(c->d_scratch64, 0, 64 * sizeof(unsigned long long)));
	*out = c;
	return 0;
}

extern "C" int cs_ctx_destroy(cs_ctx *ctx) {
	if (!ctx) return 0;
	cs_mg_destroy(ctx);
	cudaSetDevice(ctx->device);
	cudaFree(ctx->d_partials);
	cudaFree(ctx->d_partial_words);
	cudaFree(ctx->d_grid);
	if (ctx->d_remap_tab) cudaFree(ctx->d_remap_tab);
	if (ctx->d_remap_lut) cudaFree(ctx->d_remap_lut);
	cs::host_stager_destroy(ctx);
	cudaFree(ctx->d_counter);
	cudaFree(ctx->d_scratch64);
	if (ctx->d_host_buf) cudaFree(ctx->d_host_buf);
	if (ctx->host_copy) {
		cudaStreamDestroy(ctx->host_copy);
		cudaStreamDestroy(ctx->host_comp);
		for (cudaEvent_t e : ctx->host_ev) cudaEventDestroy(e);
	}
	delete ctx;
	return 0;
}

extern "C" int cs_ctx_sm_count(const cs_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

// development: copies the 64-u64 scratch block (phase stamps of CS_PHASE_TIMING builds) to the host
extern "C" int cs_debug_scratch(cs_ctx *ctx, unsigned long long *h_out64) {
	if (!ctx || !h_out64) return CS_ERR_ARG;
	CS_CUDA(cudaMemcpy(h_out64, ctx->d_scratch64, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	return 0;
}

// page-lock / release a caller-owned host buffer in place (the QImage / NumPy boundary)
extern "C" int cs_host_register(void *h_ptr, size_t bytes) {
	CS_REQUIRE(h_ptr && bytes > 0, "null pointer or empty buffer");
	CS_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable));
	return 0;
}

extern "C" int cs_host_unregister(void *h_ptr) {
	CS_REQUIRE(h_ptr, "null pointer");
	CS_CUDA(cudaHostUnregister(h_ptr));
	return 0;
}
