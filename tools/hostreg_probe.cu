/* hostreg_probe.cu - debug: growing-hull cudaHostRegister + DMA, as engine.cu: hostreg_cover does */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
struct R { uintptr_t lo, hi; };
static std::vector<R> regs;
static bool cover(const void *p, size_t bytes)
{
	const uintptr_t page = 4096;
	uintptr_t lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + bytes + page - 1) & ~(page - 1);
	for (auto &r : regs) if (lo >= r.lo && hi <= r.hi) return true;
	std::vector<R> keep;
	for (auto &r : regs) {
		if (r.hi < lo || r.lo > hi) { keep.push_back(r); continue; }
		cudaError_t e = cudaHostUnregister((void *)r.lo);
		printf("  unregister %lx..%lx: %s\n", r.lo, r.hi, cudaGetErrorString(e));
		lo = r.lo < lo ? r.lo : lo; hi = r.hi > hi ? r.hi : hi;
	}
	regs.swap(keep);
	cudaError_t e = cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterDefault);
	printf("  register %lx..%lx (%zu MiB): %s\n", lo, hi, (hi - lo) >> 20, cudaGetErrorString(e));
	if (e != cudaSuccess) { cudaGetLastError(); return false; }
	regs.push_back({lo, hi});
	return true;
}
int main()
{
	const size_t CALL = 8u << 20;
	char *buf = (char *)malloc(8 * CALL + 64);
	memset(buf, 1, 8 * CALL + 64);
	char *dev; cudaMalloc(&dev, CALL);
	cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
	for (int frame = 0; frame < 2; frame++)
		for (int c = 0; c < 8; c++) {
			const char *src = buf + 16 + c * CALL;
			bool ok = cover(src, CALL);
			cudaPointerAttributes a; cudaError_t ea = cudaPointerGetAttributes(&a, src);
			cudaError_t e = cudaMemcpyAsync(dev, src, CALL, cudaMemcpyHostToDevice, st);
			cudaError_t e2 = cudaStreamSynchronize(st);
			printf("frame %d call %d cover=%d attr=%s type=%d copy=%s sync=%s\n", frame, c, ok, cudaGetErrorString(ea), (int)a.type,
			       cudaGetErrorString(e), cudaGetErrorString(e2));
			cudaGetLastError();
		}
	return 0;
}
