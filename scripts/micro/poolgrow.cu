// How a stream-ordered memory pool grows on this box: one large allocation against many small ones, first touch
// against reuse.  Decides whether a context should size its pool in one step before the first batch.
// build: nvcc -O2 -o gpurun_variants/poolgrow scripts/micro/poolgrow.cu ; run on the GPU box
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

static double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main()
{
	cudaFree(0);
	cudaStream_t s;
	cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
	cudaMemPool_t pool;
	cudaDeviceGetDefaultMemPool(&pool, 0);
	unsigned long long thr = ~0ull;
	cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
	for (int round = 0; round < 2; round++)
	{
		// many small: 200 x 20 MB
		std::vector<void*> p(200, nullptr);
		double t0 = now_ms();
		for (auto& q : p) cudaMallocAsync(&q, 20u << 20, s);
		cudaStreamSynchronize(s);
		double t1 = now_ms();
		for (auto& q : p) cudaFreeAsync(q, s);
		cudaStreamSynchronize(s);
		printf("round %d: 200 x 20 MB cudaMallocAsync: %.2f ms\n", round, t1 - t0);
	}
	for (int round = 0; round < 2; round++)
	{
		void* q = nullptr;
		double t0 = now_ms();
		cudaError_t e = cudaMallocAsync(&q, (size_t)8 << 30, s);
		cudaStreamSynchronize(s);
		double t1 = now_ms();
		cudaFreeAsync(q, s);
		cudaStreamSynchronize(s);
		printf("round %d: 1 x 8 GB cudaMallocAsync: %.2f ms (%s)\n", round, t1 - t0, cudaGetErrorString(e));
	}
	{
		void* q = nullptr;
		double t0 = now_ms();
		cudaMalloc(&q, (size_t)8 << 30);
		double t1 = now_ms();
		cudaFree(q);
		printf("plain cudaMalloc 8 GB: %.2f ms, cudaFree %.2f ms\n", t1 - t0, now_ms() - t1);
	}
	{
		void* h = nullptr;
		double t0 = now_ms();
		cudaHostAlloc(&h, (size_t)512 << 20, cudaHostAllocDefault);
		double t1 = now_ms();
		cudaFreeHost(h);
		printf("cudaHostAlloc 512 MB: %.2f ms, cudaFreeHost %.2f ms\n", t1 - t0, now_ms() - t1);
	}
	return 0;
}
