// Integer / DPX issue-rate microbenchmark: the measured denominator of the DP roofline
// (roofline_CUPS = P_int * cells_per_vector / 6).  Each kernel runs eight
// independent dependency chains per thread of ONE instruction kind (rotating three registers
// so that nothing folds), on every SM at full occupancy, and reports warp-instructions / s.
#include "../../include/defuse_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace
{

constexpr int kChains = 8;
constexpr int kUnroll = 16; // rotations per loop iteration; 3 instructions per rotation per chain

__device__ __forceinline__ __half2 h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }
__device__ __forceinline__ uint32_t u32(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

template <int KIND>
__device__ __forceinline__ void rot(uint32_t& a, uint32_t& b, uint32_t& c)
{
	if (KIND == 0)
	{
		a = __viaddmax_s16x2(b, c, a);
		b = __viaddmax_s16x2(c, a, b);
		c = __viaddmax_s16x2(a, b, c);
	}
	else if (KIND == 1)
	{
		a = __vminu2(b, c);
		b = __vmaxu2(c, a);
		c = __vminu2(a, b) ^ 0u;
	}
	else if (KIND == 2)
	{
		a = __vimax3_s16x2(a, b, c);
		b = __vimin3_s16x2(b, c, a);
		c = __vimax3_s16x2(c, a, b);
	}
	else if (KIND == 3)
	{
		a = (a ^ b) | (c & a);
		b = (b & c) ^ (a | b);
		c = (c | a) & (b ^ c);
	}
	else if (KIND == 4)
	{
		a = a * b + c;
		b = b * c + a;
		c = c * a + b;
	}
	else if (KIND == 5)
	{
		a = a + b + c;
		b = b + c + a;
		c = c + a + b;
	}
	else if (KIND == 6)
	{
		a = __byte_perm(a, b, c);
		b = __byte_perm(b, c, a);
		c = __byte_perm(c, a, b);
	}
	else if (KIND == 8)
	{
		a = __shfl_up_sync(0xffffffffu, a, 1);
		b = __shfl_up_sync(0xffffffffu, b, 1);
		c = __shfl_up_sync(0xffffffffu, c, 1);
	}
	else if (KIND == 10) // HMNMX2
	{
		a = u32(__hmax2(h2(b), h2(c)));
		b = u32(__hmin2(h2(c), h2(a)));
		c = u32(__hmax2(h2(a), h2(b)));
	}
	else if (KIND == 11) // HADD2
	{
		a = u32(__hadd2(h2(b), h2(c)));
		b = u32(__hadd2(h2(c), h2(a)));
		c = u32(__hadd2(h2(a), h2(b)));
	}
	else if (KIND == 12) // HFMA2
	{
		a = u32(__hfma2(h2(a), h2(b), h2(c)));
		b = u32(__hfma2(h2(b), h2(c), h2(a)));
		c = u32(__hfma2(h2(c), h2(a), h2(b)));
	}
	else if (KIND == 13) // HSET2 (compare -> 1.0 / 0.0 per half)
	{
		a = u32(__hne2(h2(b), h2(c)));
		b = u32(__hne2(h2(c), h2(a)));
		c = u32(__hne2(h2(a), h2(b)));
	}
	else if (KIND == 9)
	{
		a = (uint32_t)__viaddmax_s32((int)b, (int)c, (int)a);
		b = (uint32_t)__viaddmax_s32((int)c, (int)a, (int)b);
		c = (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);
	}
}

template <int KIND>
__global__ void __launch_bounds__(256) issue_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int iters)
{
	uint32_t a[kChains], b[kChains], c[kChains];
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
	for (int k = 0; k < kChains; k++)
	{
		a[k] = in[(tid + k) & 1023];
		b[k] = in[(tid + 3 * k + 1) & 1023];
		c[k] = in[(tid + 5 * k + 2) & 1023];
	}
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int u = 0; u < kUnroll; u++)
		{
#pragma unroll
			for (int k = 0; k < kChains; k++) rot<KIND>(a[k], b[k], c[k]);
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < kChains; k++) r ^= a[k] ^ b[k] ^ c[k];
	out[tid] = r;
}

// KINDs 14..16: four chains of VIADDMNMX.S16x2 interleaved with four chains of HMNMX2 / HFMA2 / HADD2 -- do the
// two kinds issue side by side (alu + fma pipe) or queue for one pipe?
template <int HALF_KIND>
__global__ void __launch_bounds__(256) mixed_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int iters)
{
	uint32_t a[kChains], b[kChains], c[kChains];
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
	for (int k = 0; k < kChains; k++)
	{
		a[k] = in[(tid + k) & 1023];
		b[k] = in[(tid + 3 * k + 1) & 1023];
		c[k] = in[(tid + 5 * k + 2) & 1023];
	}
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int u = 0; u < kUnroll; u++)
		{
#pragma unroll
			for (int k = 0; k < kChains; k += 2)
			{
				rot<0>(a[k], b[k], c[k]);
				rot<HALF_KIND>(a[k + 1], b[k + 1], c[k + 1]);
			}
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < kChains; k++) r ^= a[k] ^ b[k] ^ c[k];
	out[tid] = r;
}

// KIND 7: the s16x2 DP cell body exactly as dp_fast_kernel issues it (S = 8 rows per lane,
// one column per iteration): LOP3, VIMNMX.U16x2, IMAD, VIADDMNMX, VIADDMNMX, VIADDMNMX per row.
__global__ void __launch_bounds__(256) cell_body_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int iters,
                                                         uint32_t xm, uint32_t g2, uint32_t gm2)
{
	constexpr int S = 8;
	uint32_t rd[S], F[S], X[S];
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
	for (int k = 0; k < S; k++)
	{
		rd[k] = in[(tid + k) & 1023] & 0x00030003u;
		F[k] = 0x40004000u;
		X[k] = 0;
	}
	uint32_t rf = in[tid & 1023] & 0x00030003u;
	uint32_t prev = 0x40004000u;
	for (int it = 0; it < iters * 6; it++)
	{
		uint32_t left = prev;
		uint32_t dg_in = prev;
		const uint32_t pen = rf & 0x80008000u;
#pragma unroll
		for (int k = 0; k < S; k++)
		{
			const uint32_t d = __vminu2(rd[k] ^ rf, 0x00010001u);
			const uint32_t dg = d * xm + dg_in;
			dg_in = F[k];
			const uint32_t e = __viaddmax_s16x2(F[k], g2, dg);
			left = __viaddmax_s16x2(left, gm2, e);
			F[k] = left;
			X[k] = __viaddmax_s16x2(left, pen, X[k]);
		}
		prev = left;
		rf = (rf + 0x00010001u) & 0x00030003u;
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < S; k++) r ^= F[k] ^ X[k];
	out[tid] = r;
}


// ---- the sweep's paired-step body in isolation (no refill, no checkpoints): what the instruction mix itself can reach ----
__device__ __forceinline__ uint32_t f16x2_add(uint32_t a, uint32_t b)
{
	uint32_t d;
	asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
	return d;
}
__device__ __forceinline__ uint32_t f16x2_fma(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__device__ __forceinline__ uint32_t f16x2_fma_sat(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

// S rows per lane, groups of 8 lanes, one SHFL.UP and one ring LDS per step like dp_fast_kernel; the last HR rows of
// the strip take their mismatch indicator from the fp16 path (HADD2, HFMA2.SAT, HFMA2: no ALU-pipe issue), the others
// from VIADDMNMX.U16x2 + IMAD.  SINK: 0 none (probe-like), 1 = one VIMNMX3 per row per two steps (split / simple sweeps)
template <int S, int HR, int SINK>
__global__ void __launch_bounds__(128, 4) sweep_body_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int pairs,
                                                             uint32_t xm, uint32_t g2, uint32_t gm2, uint32_t c2)
{
	__shared__ uint32_t s_ring[4][4][128 + 8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane >> 3, g = lane & 7;
	uint32_t* ring = s_ring[warp][q];
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	for (int k = g; k < 128; k += 8)
	{
		const uint32_t a = in[(tid + 7 * k) & 1023], b = in[(tid + 11 * k + 5) & 1023];
		const uint32_t sym[4] = {0x5410u, 0x5430u, 0x5470u, 0x5540u}; // fp16 bit patterns of 65, 67, 71, 84
		ring[k] = sym[a & 3] | (sym[b & 3] << 16);
	}
	__syncwarp();
	uint32_t rd[S], F[S], Fb[S], Y[S];
#pragma unroll
	for (int k = 0; k < S; k++)
	{
		const uint32_t sy = ring[(k * 5 + g) & 127];
		rd[k] = (k >= S - HR) ? (sy ^ 0x80008000u) : ((0u - (sy & 0xFFFFu)) & 0xFFFFu) | ((0u - (sy >> 16)) << 16);
		F[k] = 0x01900190u + 0x00010001u * (uint32_t)k;
		Y[k] = 0;
	}
	uint32_t prev = 0x01900190u, Flast = F[S - 1];
	const uint32_t Bp = 0x01A001A0u;
	auto half_step = [&](const uint32_t* Fin, uint32_t* Fout, const int u, const bool odd) {
		uint32_t recv = __shfl_up_sync(0xffffffffu, Flast, 1, 8);
		if (g == 0) recv = Bp;
		const uint32_t rf = ring[(u - g) & 127];
		uint32_t left = recv, dg_in = prev;
#pragma unroll
		for (int k = 0; k < S; k++)
		{
			uint32_t dg;
			if (k >= S - HR)
			{
				const uint32_t d = f16x2_add(rf, rd[k]);
				const uint32_t r = f16x2_fma_sat(d, d, 0u);
				dg = f16x2_fma(r, c2, dg_in);
			}
			else
			{
				const uint32_t d = __viaddmin_u16x2(rd[k], rf, 0x00010001u);
				dg = d * xm + dg_in;
			}
			const uint32_t fold = Fin[k];
			dg_in = fold;
			const uint32_t e = __viaddmax_s16x2(fold, g2, dg);
			left = __viaddmax_s16x2(left, gm2, e);
			Fout[k] = left;
			if (SINK == 1 && odd) Y[k] = __vimax3_s16x2(Y[k], fold, left);
		}
		prev = recv;
		Flast = left;
	};
#pragma unroll 1
	for (int it = 0, u = 8; it < pairs; it++, u += 2)
	{
		half_step(F, Fb, u, false);
		half_step(Fb, F, u + 1, true);
	}
	uint32_t r = 0;
#pragma unroll
	for (int k = 0; k < S; k++) r ^= F[k] ^ Y[k];
	out[tid] = r;
}

template <int S, int HR, int SINK>
static void launch_body(int sm_count, const uint32_t* d_in, uint32_t* d_out, int pairs)
{
	// 2/-1/-2: xm = x - m = -3 as a 32-bit multiplier, gap both halves, gap - match both halves, -(m - x) as an fp16 subnormal
	sweep_body_kernel<S, HR, SINK><<<sm_count * 4, 128>>>(d_in, d_out, pairs, 0xFFFFFFFDu, 0xFFFEFFFEu, 0xFFFCFFFCu, 0x80038003u);
}

}  // namespace

extern "C" int dfb_microbench_issue_rate(dfb_ctx* ctx, int kind, int iters, double* warp_instr_per_s, double* elapsed_ms)
{
	// kinds 0..16: one instruction kind (or a 1:1 mix) in independent chains, result = warp-instructions / s.
	// kinds 100+HR / 200+HR: the sweep's paired-step body with S = 13 rows per lane, HR of them on the fp16 indicator
	// path, with (100) / without (200) the row-maximum sink; `iters` = step pairs, result = warp row-steps / s
	// (one row-step = one register pair of cells = 64 cell updates per warp)
	const bool body = kind >= 100;
	const int body_hr = body ? kind % 100 : 0;
	if (!ctx || !warp_instr_per_s || kind < 0 || (!body && kind > 16) || (body && (kind >= 300 || body_hr > 13)) || iters <= 0)
		return DFB_ERR_ARG;
	dfb_device_info info;
	int rc = dfb_ctx_device_info(ctx, &info);
	if (rc) return rc;
	if (cudaSetDevice(info.ordinal) != cudaSuccess) return DFB_ERR_CUDA;
	const int blocks = body ? info.sm_count * 4 : info.sm_count * 8;
	const int threads = body ? 128 : 256;
	uint32_t* d_in = nullptr;
	uint32_t* d_out = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	auto cleanup = [&]() {
		if (e0) cudaEventDestroy(e0);
		if (e1) cudaEventDestroy(e1);
		cudaFree(d_in);
		cudaFree(d_out);
	};
	if (cudaMalloc(&d_in, 1024 * sizeof(uint32_t)) != cudaSuccess) return DFB_ERR_NOMEM;
	if (cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(uint32_t)) != cudaSuccess)
	{
		cleanup();
		return DFB_ERR_NOMEM;
	}
	uint32_t h_in[1024];
	uint32_t s = 12345u;
	for (int k = 0; k < 1024; k++)
	{
		s = s * 1664525u + 1013904223u;
		h_in[k] = (s >> 4) & 0x0FFF0FFFu;
	}
	if (cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess ||
	    cudaEventCreate(&e1) != cudaSuccess)
	{
		cleanup();
		return DFB_ERR_CUDA;
	}
	float best_ms = 1e30f;
	bool known = true;
	for (int rep = 0; rep < 4 && known; rep++) // first repetition is the warm-up
	{
		cudaEventRecord(e0, 0);
		switch (kind)
		{
			case 0: issue_kernel<0><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 1: issue_kernel<1><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 2: issue_kernel<2><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 3: issue_kernel<3><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 4: issue_kernel<4><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 5: issue_kernel<5><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 6: issue_kernel<6><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 7: cell_body_kernel<<<blocks, threads>>>(d_in, d_out, iters, 0xFFFFFFFDu, 0xFFFEFFFEu, 0xFFFCFFFCu); break;
			case 8: issue_kernel<8><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 9: issue_kernel<9><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 10: issue_kernel<10><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 11: issue_kernel<11><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 12: issue_kernel<12><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 13: issue_kernel<13><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 14: mixed_kernel<10><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 15: mixed_kernel<12><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 16: mixed_kernel<11><<<blocks, threads>>>(d_in, d_out, iters); break;
			case 100: launch_body<13, 0, 1>(info.sm_count, d_in, d_out, iters); break;
			case 103: launch_body<13, 3, 1>(info.sm_count, d_in, d_out, iters); break;
			case 105: launch_body<13, 5, 1>(info.sm_count, d_in, d_out, iters); break;
			case 107: launch_body<13, 7, 1>(info.sm_count, d_in, d_out, iters); break;
			case 109: launch_body<13, 9, 1>(info.sm_count, d_in, d_out, iters); break;
			case 111: launch_body<13, 11, 1>(info.sm_count, d_in, d_out, iters); break;
			case 113: launch_body<13, 13, 1>(info.sm_count, d_in, d_out, iters); break;
			case 200: launch_body<13, 0, 0>(info.sm_count, d_in, d_out, iters); break;
			case 207: launch_body<13, 7, 0>(info.sm_count, d_in, d_out, iters); break;
			case 213: launch_body<13, 13, 0>(info.sm_count, d_in, d_out, iters); break;
			default: known = false; break;
		}
		if (!known) break;
		cudaEventRecord(e1, 0);
		if (cudaEventSynchronize(e1) != cudaSuccess) break;
		float ms = 0;
		if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) break;
		if (rep > 0 && ms < best_ms) best_ms = ms;
	}
	cudaError_t e = cudaDeviceSynchronize();
	cleanup();
	if (!known) return DFB_ERR_ARG;
	if (e != cudaSuccess || best_ms > 1e29f) return DFB_ERR_CUDA;
	const double warps = (double)blocks * threads / 32.0;
	// kind 7 counts one "body" = the 6 instructions of one register-pair of cells; the sweep bodies count row-steps
	const double per_thread = body ? (double)iters * 2 * 13 : (kind == 7) ? (double)iters * 6 * 8 : (double)iters * kUnroll * kChains * 3;
	*warp_instr_per_s = warps * per_thread / (best_ms * 1e-3);
	if (elapsed_ms) *elapsed_ms = best_ms;
	return DFB_OK;
}
