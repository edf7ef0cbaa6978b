// Result assembly on the GPU: the second half of SplitReadAligner::GetAlignments (tools/SplitReadAligner.cpp:229-285)
// for the common task -- at most DFB_SLOT_EVENTS arg-max columns, all of them in the task's fixed 128-byte region
// left by the probe sweep.  The kernels order those columns the way GetAlignments walks them (matrix, row, column),
// pair every winning row a of matrix 1 with row L-a of matrix 2 and write the dfb_split_row records and their column
// lists in TASK ORDER into two dense arrays, so that the host only copies them back.  Tasks whose columns spilled into
// the overflow list (tie-heavy ones; a fraction of a percent of a dosplitalign batch) are left to the host: their region entries
// are appended to that list here, which makes the list self-contained.
//
//   asm_count_kernel   rows / columns per task -> sums per block of 256 tasks
//   asm_scan_kernel    exclusive scan of the block sums (one block), totals
//   asm_write_kernel   per-task offsets by an in-block scan, records written
// HBM-bound and tiny next to the sweeps: 128 B read and <= 320 B written per winning task.
#pragma once

namespace dfb
{
#define DFB_ASM_MAX_CLASSES 16
#define DFB_ASM_BLOCK 256

struct AsmParams
{
	long long n_tasks;
	const int32_t* task_slot; // global slot per task, -1: none
	int n_classes;
	long long job_base[DFB_ASM_MAX_CLASSES + 1]; // first global slot of every class (slots are numbered in job space)
	const int* slot_n[DFB_ASM_MAX_CLASSES];
	const uint2* slot_ev[DFB_ASM_MAX_CLASSES];
	const int* hitq[DFB_ASM_MAX_CLASSES];
	const JobPair* jobs[DFB_ASM_MAX_CLASSES];
	unsigned long long* block_sums; // [n_blocks][2] rows, cols; scanned in place into exclusive offsets
	unsigned long long* totals;     // [2]
	dfb_split_row* rows;
	int32_t* cols;
	Event* events; // overflow list (append)
	unsigned long long* ev_count;
	unsigned long long ev_cap;
	unsigned long long rows_cap, cols_cap; // entries allocated (bounds-check builds)
};

// the task's region entries ordered by key (matrix << 27 | row << 16 | column); returns their number, -1 for a task
// that has none to assemble here (no slot, or spilled)
__device__ __forceinline__ int asm_load(const AsmParams& p, long long t, uint2 (&e)[DFB_SLOT_EVENTS], int& L, int& cls, long long& local)
{
	const int32_t s = p.task_slot[t];
	if (s < 0) return -1;
	int c = 0;
	while (c + 1 < p.n_classes && (long long)s >= p.job_base[c + 1]) c++;
	cls = c;
	local = (long long)s - p.job_base[c];
	const int n = p.slot_n[c][local];
	if (n > DFB_SLOT_EVENTS) return -2;
	L = p.jobs[c][p.hitq[c][local]].L[0];
	const uint2* src = p.slot_ev[c] + local * DFB_SLOT_EVENTS;
	for (int k = 0; k < n; k++)
	{
		const uint2 v = src[k];
		int b = k - 1;
		while (b >= 0 && e[b].x > v.x)
		{
			e[b + 1] = e[b];
			b--;
		}
		e[b + 1] = v;
	}
	return n;
}

// walks the rows of one task; F(a, i, i_end, j, j_end) is called for every emitted row
template <class F>
__device__ __forceinline__ void asm_rows(const uint2 (&e)[DFB_SLOT_EVENTS], int n, int L, F emit)
{
	int n0 = 0;
	while (n0 < n && !(e[n0].x >> 27)) n0++;
	int i = 0;
	while (i < n0)
	{
		const uint32_t a = (e[i].x >> 16) & 0x7ff;
		int i_end = i + 1;
		while (i_end < n0 && ((e[i_end].x >> 16) & 0x7ff) == a) i_end++;
		const uint32_t want = (uint32_t)L - a;
		int j = n0;
		while (j < n && ((e[j].x >> 16) & 0x7ff) != want) j++;
		if (j < n)
		{
			int j_end = j + 1;
			while (j_end < n && ((e[j_end].x >> 16) & 0x7ff) == want) j_end++;
			emit((int)a, i, i_end, j, j_end);
		}
		i = i_end;
	}
}

__global__ void __launch_bounds__(DFB_ASM_BLOCK) asm_count_kernel(AsmParams p)
{
	__shared__ unsigned int s_rows, s_cols;
	if (threadIdx.x == 0) s_rows = s_cols = 0;
	__syncthreads();
	const long long t = (long long)blockIdx.x * DFB_ASM_BLOCK + threadIdx.x;
	unsigned nr = 0, nc = 0;
	if (t < p.n_tasks)
	{
		uint2 e[DFB_SLOT_EVENTS];
		int L = 0, cls = 0;
		long long local = 0;
		const int n = asm_load(p, t, e, L, cls, local);
		if (n == -2)
		{
			// spilled: hand the region's entries to the overflow list, the host assembles the task from that list alone
			const uint2* src = p.slot_ev[cls] + local * DFB_SLOT_EVENTS;
			for (int k = 0; k < DFB_SLOT_EVENTS; k++)
			{
				const unsigned long long idx = atomicAdd(p.ev_count, 1ull);
				if (idx < p.ev_cap)
				{
					const uint2 v = src[k];
					Event ev;
					ev.task = (int32_t)t;
					ev.half_row = (int32_t)(((v.x >> 27) << 30) | ((v.x >> 16) & 0x7ff));
					ev.col = (int32_t)(v.x & 0xffff);
					ev.score = (int32_t)v.y;
					p.events[idx] = ev;
				}
			}
		}
		else if (n > 0)
		{
			asm_rows(e, n, L, [&](int, int i, int i_end, int j, int j_end) {
				nr++;
				nc += (unsigned)((i_end - i) + (j_end - j));
			});
		}
	}
	// block sums (warp reduce, then one atomic per warp)
	for (int o = 16; o > 0; o >>= 1)
	{
		nr += __shfl_down_sync(0xffffffffu, nr, o);
		nc += __shfl_down_sync(0xffffffffu, nc, o);
	}
	if ((threadIdx.x & 31) == 0)
	{
		atomicAdd(&s_rows, nr);
		atomicAdd(&s_cols, nc);
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		p.block_sums[2 * (size_t)blockIdx.x] = s_rows;
		p.block_sums[2 * (size_t)blockIdx.x + 1] = s_cols;
	}
}

// one block: exclusive scan of the per-block sums, in place
__global__ void __launch_bounds__(1024) asm_scan_kernel(unsigned long long* sums, long long n_blocks, unsigned long long* totals)
{
	__shared__ unsigned long long warp_r[32], warp_c[32];
	__shared__ unsigned long long carry_r, carry_c;
	if (threadIdx.x == 0) carry_r = carry_c = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (long long base = 0; base < n_blocks; base += 1024)
	{
		const long long k = base + threadIdx.x;
		unsigned long long r = k < n_blocks ? sums[2 * k] : 0, c = k < n_blocks ? sums[2 * k + 1] : 0;
		const unsigned long long r0 = r, c0 = c;
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long ur = __shfl_up_sync(0xffffffffu, r, o), uc = __shfl_up_sync(0xffffffffu, c, o);
			if (lane >= o)
			{
				r += ur;
				c += uc;
			}
		}
		if (lane == 31)
		{
			warp_r[warp] = r;
			warp_c[warp] = c;
		}
		__syncthreads();
		if (warp == 0)
		{
			unsigned long long wr = warp_r[lane], wc = warp_c[lane];
			for (int o = 1; o < 32; o <<= 1)
			{
				const unsigned long long ur = __shfl_up_sync(0xffffffffu, wr, o), uc = __shfl_up_sync(0xffffffffu, wc, o);
				if (lane >= o)
				{
					wr += ur;
					wc += uc;
				}
			}
			warp_r[lane] = wr; // inclusive over warps
			warp_c[lane] = wc;
		}
		__syncthreads();
		const unsigned long long before_r = carry_r + (warp ? warp_r[warp - 1] : 0), before_c = carry_c + (warp ? warp_c[warp - 1] : 0);
		if (k < n_blocks)
		{
			sums[2 * k] = before_r + r - r0;
			sums[2 * k + 1] = before_c + c - c0;
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			carry_r += warp_r[31];
			carry_c += warp_c[31];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		totals[0] = carry_r;
		totals[1] = carry_c;
	}
}

__global__ void __launch_bounds__(DFB_ASM_BLOCK) asm_write_kernel(AsmParams p)
{
	__shared__ unsigned int w_rows[DFB_ASM_BLOCK / 32], w_cols[DFB_ASM_BLOCK / 32];
	const long long t = (long long)blockIdx.x * DFB_ASM_BLOCK + threadIdx.x;
	uint2 e[DFB_SLOT_EVENTS];
	int L = 0, cls = 0, n = -1;
	long long local = 0;
	unsigned nr = 0, nc = 0;
	if (t < p.n_tasks)
	{
		n = asm_load(p, t, e, L, cls, local);
		if (n > 0)
			asm_rows(e, n, L, [&](int, int i, int i_end, int j, int j_end) {
				nr++;
				nc += (unsigned)((i_end - i) + (j_end - j));
			});
	}
	// exclusive offsets inside the block
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned ir = nr, ic = nc;
	for (int o = 1; o < 32; o <<= 1)
	{
		const unsigned ur = __shfl_up_sync(0xffffffffu, ir, o), uc = __shfl_up_sync(0xffffffffu, ic, o);
		if (lane >= o)
		{
			ir += ur;
			ic += uc;
		}
	}
	if (lane == 31)
	{
		w_rows[warp] = ir;
		w_cols[warp] = ic;
	}
	__syncthreads();
	unsigned long long row_at = p.block_sums[2 * (size_t)blockIdx.x] + (ir - nr);
	unsigned long long col_at = p.block_sums[2 * (size_t)blockIdx.x + 1] + (ic - nc);
	for (int w = 0; w < warp; w++)
	{
		row_at += w_rows[w];
		col_at += w_cols[w];
	}
	if (n > 0 && nr)
	{
		asm_rows(e, n, L, [&](int a, int i, int i_end, int j, int j_end) {
			dfb_split_row row;
			row.task = (int32_t)t;
			row.read_split = a;
			row.score1 = (int32_t)e[i].y;
			row.score2 = (int32_t)e[j].y;
			row.col_begin = (int64_t)col_at;
			row.n1 = i_end - i;
			row.n2 = j_end - j;
			DFB_BC(row_at < p.rows_cap && col_at + (unsigned)(row.n1 + row.n2) <= p.cols_cap, 501);
			p.rows[row_at++] = row;
			for (int k = i; k < i_end; k++) p.cols[col_at++] = (int32_t)(e[k].x & 0xffff);
			for (int k = j; k < j_end; k++) p.cols[col_at++] = (int32_t)(e[k].x & 0xffff);
		});
	}
}
}  // namespace dfb
