// hostio.cu — staged upload of PAGEABLE caller memory (after the kernels the copies dominate).
//
// The reference hands images over as ordinary NumPy arrays (app/ui/main_window.py:596-601: the colour panel passes
// `self.working_image` to color_simplify.*).  A cudaMemcpy from pageable memory is staged by the driver through
// its own bounce buffers on ONE host thread (~10-12 GB/s measured here: 268 MB of RGBA in ~22 ms), page-locking the
// caller's buffer in place (cs_host_register) costs as much as the copy it saves when the array is used once.
// cs_host_upload does the staging itself: a small pool of host threads copies the array, 8 MB at a time, into a ring
// of page-locked buffers, and every filled buffer leaves as one asynchronous DMA on the caller's stream while the
// threads fill the next one.  No reference code is involved: this is plumbing at the library's boundary.
#include "cs_common.cuh"

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace cs {

struct HostStager {
	static constexpr int kBufs = 4;
	static constexpr size_t kChunk = (size_t)8 << 20;  // bytes per DMA
	static constexpr size_t kPart = (size_t)512 << 10;  // bytes per memcpy work item
	void *buf[kBufs] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev[kBufs];
	bool used[kBufs] = {false, false, false, false};
	// worker pool: one job at a time = copy `bytes` from src to dst in kPart pieces
	std::vector<std::thread> workers;
	std::mutex m;
	std::condition_variable cv_work;
	const char *src = nullptr;
	char *dst = nullptr;
	size_t bytes = 0;
	std::atomic<long long> next_part{0}, done_parts{0};
	long long nparts = 0;
	unsigned long long gen = 0;
	bool stop = false;

	std::atomic<int> active{0};  // workers inside run_parts (incremented under `m`, so the owner's check below is safe)

	struct Job {
		const char *src;
		char *dst;
		size_t bytes;
		long long nparts;
	};
	void run_parts(const Job &j) {
		for (;;) {
			const long long p = next_part.fetch_add(1, std::memory_order_relaxed);
			if (p >= j.nparts) break;
			const size_t o = (size_t)p * kPart, len = std::min(kPart, j.bytes - o);
			std::memcpy(j.dst + o, j.src + o, len);
			done_parts.fetch_add(1, std::memory_order_release);
		}
	}
	void worker() {
		unsigned long long seen = 0;
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> lk(m);
				cv_work.wait(lk, [&] { return stop || gen != seen; });
				if (stop) return;
				seen = gen;
				j = Job{src, dst, bytes, nparts};
				active.fetch_add(1, std::memory_order_relaxed);
			}
			run_parts(j);
			active.fetch_sub(1, std::memory_order_release);
		}
	}
	void parallel_copy(void *d, const void *s, size_t n) {
		Job j{static_cast<const char *>(s), static_cast<char *>(d), n, (long long)((n + kPart - 1) / kPart)};
		for (;;) {  // a new job is published only when no worker is still looking at the previous one
			std::unique_lock<std::mutex> lk(m);
			if (active.load(std::memory_order_acquire) == 0) {
				src = j.src; dst = j.dst; bytes = j.bytes; nparts = j.nparts;
				next_part.store(0, std::memory_order_relaxed);
				done_parts.store(0, std::memory_order_relaxed);
				++gen;
				break;
			}
			lk.unlock();
			std::this_thread::yield();
		}
		cv_work.notify_all();
		run_parts(j);  // the calling thread works too
		while (done_parts.load(std::memory_order_acquire) < j.nparts) std::this_thread::yield();
	}
};

static int stager_get(cs_ctx *ctx, HostStager **out) {
	if (!ctx->host_stager) {
		HostStager *h = new HostStager();
		for (int i = 0; i < HostStager::kBufs; ++i) {
			CS_CUDA(cudaHostAlloc(&h->buf[i], HostStager::kChunk, cudaHostAllocPortable));
			CS_CUDA(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
		}
		int nt = 0;
		if (const char *e = getenv("CS_HOST_THREADS")) nt = atoi(e);
		if (nt <= 0) {
			// the CPUs this process may run on (measured on the 16-CPU host of a B200 box, 268 MB: 2 threads 18 GB/s,
			// 4: 33, 8: 45, 12: 46, 16: 43)
			cpu_set_t set;
			int ncpu = (int)std::thread::hardware_concurrency();
			if (sched_getaffinity(0, sizeof(set), &set) == 0) ncpu = CPU_COUNT(&set);
			nt = std::min(12, std::max(2, (ncpu * 3) / 4));
		}
		for (int i = 0; i < nt - 1; ++i) h->workers.emplace_back([h] { h->worker(); });
		ctx->host_stager = h;
	}
	*out = static_cast<HostStager *>(ctx->host_stager);
	return 0;
}

void host_stager_destroy(cs_ctx *ctx) {
	HostStager *h = static_cast<HostStager *>(ctx->host_stager);
	if (!h) return;
	{
		std::lock_guard<std::mutex> lk(h->m);
		h->stop = true;
	}
	h->cv_work.notify_all();
	for (std::thread &t : h->workers) t.join();
	for (int i = 0; i < HostStager::kBufs; ++i) {
		if (h->used[i]) cudaEventSynchronize(h->ev[i]);
		if (h->buf[i]) cudaFreeHost(h->buf[i]);
		cudaEventDestroy(h->ev[i]);
	}
	delete h;
	ctx->host_stager = nullptr;
}

} // namespace cs

using namespace cs;

extern "C" int cs_host_upload(cs_ctx *ctx, const void *h_src, size_t bytes, void *d_dst, void *stream) {
	CS_REQUIRE(ctx && (bytes == 0 || (h_src && d_dst)), "null pointer");
	if (bytes == 0) return 0;
	CS_CUDA(cudaSetDevice(ctx->device));
	HostStager *h = nullptr;
	const int rc = stager_get(ctx, &h);
	if (rc) return rc;
	cudaStream_t st = (cudaStream_t)stream;
	int i = 0;
	for (size_t off = 0; off < bytes; off += HostStager::kChunk, ++i) {
		const int b = i % HostStager::kBufs;
		const size_t len = std::min(HostStager::kChunk, bytes - off);
		if (h->used[b]) CS_CUDA(cudaEventSynchronize(h->ev[b]));  // its previous DMA has read the buffer
		h->parallel_copy(h->buf[b], static_cast<const char *>(h_src) + off, len);
		CS_CUDA(cudaMemcpyAsync(static_cast<char *>(d_dst) + off, h->buf[b], len, cudaMemcpyHostToDevice, st));
		CS_CUDA(cudaEventRecord(h->ev[b], st));
		h->used[b] = true;
	}
	return 0;
}
