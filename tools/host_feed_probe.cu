/*
 * host_feed_probe.cu - what bounds the host-fed path on the GPU box (not part of the product):
 *   - CPU memcpy pageable -> page-locked with 1..16 threads (the staging copy of pageable callers)
 *   - cudaMemcpyAsync straight from pageable memory (driver staging)
 *   - cudaHostRegister / Unregister cost per MiB
 *   - page-locked H2D at 8 MiB
 * Build: nvcc -O3 -std=c++17 -o /tmp/host_feed_probe tools/host_feed_probe.cu -lpthread
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main()
{
	const size_t CALL = 8u << 20;              /* one fosphor_cl_process call: 1024 x 1024 cf32 */
	const size_t POOL = 1u << 30;              /* pageable source pool >> LLC */
	char *src = (char *)malloc(POOL);
	memset(src, 1, POOL);
	char *pin = nullptr, *dev = nullptr;
	CK(cudaMallocHost(&pin, 4 * CALL));
	CK(cudaMalloc(&dev, 4 * CALL));
	cudaStream_t st;
	CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());

	/* 1. threaded memcpy pageable -> pinned, 8 MiB per "call", fresh threads per call are too slow:
	 *    measure with persistent threads spinning on a flag */
	for (int T : {1, 2, 4, 8, 12, 16}) {
		std::vector<std::thread> th;
		volatile int go = 0;
		volatile int stop = 0;
		volatile int done[16] = {0};
		const char *cur = src;
		for (int t = 0; t < T; t++)
			th.emplace_back([&, t] {
				int seen = 0;
				while (!stop) {
					if (go == seen) { continue; }
					seen = go;
					const size_t per = CALL / T;
					memcpy(pin + t * per, cur + t * per, per);
					done[t] = seen;
				}
			});
		const int reps = 100;
		double t0 = now();
		for (int r = 1; r <= reps; r++) {
			cur = src + ((size_t)r * CALL) % (POOL - CALL);
			__sync_synchronize();
			go = r;
			for (int t = 0; t < T; t++)
				while (done[t] != r) { }
		}
		double el = now() - t0;
		stop = 1;
		go = -1;
		for (auto &x : th) x.join();
		printf("memcpy pageable->pinned  %2d threads: %6.1f GB/s  (%.0f us per 8 MiB)\n", T, reps * (double)CALL / el / 1e9, el / reps * 1e6);
	}

	/* 2. driver-staged pageable H2D */
	{
		const int reps = 50;
		CK(cudaStreamSynchronize(st));
		double t0 = now();
		for (int r = 0; r < reps; r++)
			CK(cudaMemcpyAsync(dev, src + ((size_t)r * CALL) % (POOL - CALL), CALL, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		double el = now() - t0;
		printf("cudaMemcpyAsync from pageable: %6.1f GB/s (%.0f us per 8 MiB)\n", reps * (double)CALL / el / 1e9, el / reps * 1e6);
	}
	/* 3. pinned H2D */
	{
		const int reps = 200;
		double t0 = now();
		for (int r = 0; r < reps; r++)
			CK(cudaMemcpyAsync(dev + (r & 3) * CALL, pin + (r & 3) * CALL, CALL, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		double el = now() - t0;
		printf("cudaMemcpyAsync from pinned:   %6.1f GB/s (%.0f us per 8 MiB)\n", reps * (double)CALL / el / 1e9, el / reps * 1e6);
		/* 1 MiB pieces */
		t0 = now();
		for (int r = 0; r < reps * 8; r++)
			CK(cudaMemcpyAsync(dev + (r & 31) * (CALL / 8), pin + (r & 31) * (CALL / 8), CALL / 8, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		el = now() - t0;
		printf("  ... in 1 MiB pieces:         %6.1f GB/s\n", reps * (double)CALL / el / 1e9);
		/* D2H 4.5 MiB to pinned / to pageable */
		t0 = now();
		for (int r = 0; r < reps; r++)
			CK(cudaMemcpyAsync(pin, dev, 4718592, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		el = now() - t0;
		printf("D2H 4.5 MiB to pinned:   %6.1f GB/s (%.0f us)\n", reps * 4718592.0 / el / 1e9, el / reps * 1e6);
		t0 = now();
		for (int r = 0; r < 50; r++)
			CK(cudaMemcpyAsync(src + (size_t)r * CALL, dev, 4718592, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		el = now() - t0;
		printf("D2H 4.5 MiB to pageable: %6.1f GB/s (%.0f us)\n", 50 * 4718592.0 / el / 1e9, el / 50 * 1e6);
	}
	/* 4. register / unregister */
	for (size_t mb : {8, 16, 64}) {
		char *p = src + (64u << 20);
		double t0 = now();
		cudaError_t e = cudaHostRegister(p, mb << 20, cudaHostRegisterDefault);
		double t1 = now();
		if (e != cudaSuccess) { printf("cudaHostRegister failed: %s\n", cudaGetErrorString(e)); break; }
		CK(cudaMemcpyAsync(dev, p, CALL, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		double t2 = now();
		for (int r = 0; r < 20; r++)
			CK(cudaMemcpyAsync(dev, p, CALL, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		double t3 = now();
		CK(cudaHostUnregister(p));
		double t4 = now();
		printf("cudaHostRegister %3zu MiB: %.0f us (%.1f GB/s), first copy %.0f us, steady H2D %.1f GB/s, unregister %.0f us\n",
		       mb, (t1 - t0) * 1e6, (double)(mb << 20) / (t1 - t0) / 1e9, (t2 - t1) * 1e6, 20.0 * CALL / (t3 - t2) / 1e9, (t4 - t3) * 1e6);
	}
	return 0;
}
