// What starting CUDA costs a process on this box with nothing of ours in it: the floor under every tool's wall clock.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/cudainit scripts/micro/cudainit.cu
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

static double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
__global__ void k(int* p) { *p = 1; }

int main()
{
	double t0 = now_ms();
	int n = 0;
	cudaGetDeviceCount(&n);
	double t1 = now_ms();
	cudaSetDevice(0);
	cudaFree(0);
	double t2 = now_ms();
	cudaStream_t s;
	cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
	int* d = nullptr;
	cudaMalloc(&d, 4);
	k<<<1, 1, 0, s>>>(d);
	cudaStreamSynchronize(s);
	double t3 = now_ms();
	printf("devices %d: cudaGetDeviceCount %.1f ms, context (cudaSetDevice + cudaFree(0)) %.1f ms, stream + first launch %.1f ms\n", n, t1 - t0, t2 - t1,
	       t3 - t2);
	return 0;
}
