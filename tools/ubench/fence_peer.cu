// Does a gpu-scope fence behind a stream of ordinary stores get more expensive once peer access is enabled?
// (development probe for DESIGN.md open item 3; needs >= 2 GPUs: `gpurun --gpus 2 -- tools/ubench/fence_peer`)
//
// One persistent CTA per SM writes `bytes_per_cta` of its own device memory with 32-bit stores (like the label
// store of the Lloyd kernel), then thread 0 executes __threadfence() + atomicAdd on a counter (the "blocks
// finished" election).  Reported: kernel time and, from %globaltimer, the time thread 0 of each CTA spends
// between its last store and the return of the atomic, without peer access and with peer access to 1..N-1 peers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gt() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

template <int MODE>  // 0 = fence + atomic, 1 = atomic only (no fence)
__global__ void __launch_bounds__(512, 1) k(uint32_t *dst, size_t words_per_cta, unsigned int *counter,
                                             unsigned long long *t_fence) {
	uint32_t *mine = dst + (size_t)blockIdx.x * words_per_cta;
	for (size_t i = threadIdx.x; i < words_per_cta; i += blockDim.x) mine[i] = (uint32_t)i;
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned long long t0 = gt();
		if (MODE == 0) __threadfence();
		atomicAdd(counter, 1u);
		t_fence[blockIdx.x] = gt() - t0;
	}
}

template <int MODE> void run(const char *what, int dev, size_t bytes_per_cta) {
	CK(cudaSetDevice(dev));
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
	const int grid = p.multiProcessorCount;
	const size_t words = bytes_per_cta / 4;
	uint32_t *dst; unsigned int *ctr; unsigned long long *tf;
	CK(cudaMalloc(&dst, (size_t)grid * words * 4)); CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&tf, grid * 8));
	CK(cudaMemset(ctr, 0, 4));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int i = 0; i < 5; ++i) k<MODE><<<grid, 512>>>(dst, words, ctr, tf);
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(e0));
	const int reps = 50;
	for (int i = 0; i < reps; ++i) k<MODE><<<grid, 512>>>(dst, words, ctr, tf);
	CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	std::vector<unsigned long long> h(grid);
	CK(cudaMemcpy(h.data(), tf, grid * 8, cudaMemcpyDeviceToHost));
	double avg = 0; unsigned long long mx = 0;
	for (auto v : h) { avg += (double)v; if (v > mx) mx = v; }
	printf("%-44s %6.1f MB/launch: %8.2f us per launch, fence+atomic per CTA avg %7.0f ns max %7llu ns\n", what,
	       (double)grid * words * 4 / 1e6, ms * 1e3 / reps, avg / grid, mx);
	CK(cudaFree(dst)); CK(cudaFree(ctr)); CK(cudaFree(tf));
}

int main() {
	int n = 0; CK(cudaGetDeviceCount(&n));
	printf("%d GPUs\n", n);
	const size_t sizes[] = {4096, 453 * 1024, 4 * 453 * 1024};  // 453 KB = one CTA's share of a 64 MP label map
	for (size_t b : sizes) { run<0>("no peer access, fence + atomic", 0, b); run<1>("no peer access, atomic only", 0, b); }
	for (int peers = 1; peers < n; ++peers) {
		CK(cudaSetDevice(0));
		int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, peers));
		if (!can) { printf("no peer access 0 -> %d\n", peers); break; }
		CK(cudaDeviceEnablePeerAccess(peers, 0));
		CK(cudaSetDevice(peers)); CK(cudaDeviceEnablePeerAccess(0, 0));
		char name[96];
		for (size_t b : sizes) {
			snprintf(name, sizeof name, "%d peer(s) enabled, fence + atomic", peers); run<0>(name, 0, b);
			snprintf(name, sizeof name, "%d peer(s) enabled, atomic only", peers); run<1>(name, 0, b);
		}
	}
	return 0;
}
