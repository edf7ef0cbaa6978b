// Backtrace of chosen split alignments: SplitReadAligner::GetAlignments(..., backtrace=true)
// (tools/SplitReadAligner.cpp:56-69 pointer rule, :124-143 BackTracePath, :145-154 ReverseMatches, :287-292).
//
// Included at the end of dfb_api.cu (it needs dfb_ctx).  The workload is small by construction -- splitseq
// re-aligns only the reads that support one predicted break (tools/splitseq.cpp:100-125) -- so this path is
// built for exactness over every parameter set (s32 arithmetic, raw byte equality, any length), not for TCUPS:
//
//   trace_fill_kernel   one warp per matrix (a task has two: read x ref1, rev(read) x rev(ref2)).  Lane l owns
//                       read row j = tile + l + 1 and sweeps the reference as a skewed wavefront, exactly the
//                       geometry of the s32 DP kernel; what is new is the pointer of every cell, 2 bits, packed
//                       16 columns per word by the lane that owns the row:
//                         0 diagonal, 1 (i-1,j), 2 (i,j-1); later writers win, so 2 > 1 > 0 (:56-69)
//   trace_walk_kernel   one thread per matrix walks the pointers from the start cell until j == 0 and records
//                       the diagonal steps in walk order (descending positions).
// The host reverses matrix 1's list and maps matrix 2's coordinates back to the unreversed sequences (whose
// double reversal is the walk order itself).
#pragma once

namespace dfb
{
struct TraceJob
{
	long long ref_off;  // first byte of the reference in the raw pool
	long long read_off; // first byte of the read in the raw pool
	int R, L;
	int reversed;       // 1: both sequences are consumed back to front (matrix 2)
	int start_i, start_j;
	long long bits_off;    // words; (L+1) rows of W = (R+16)/16 words
	long long scratch_off; // ints; R+1 row-boundary values
	long long match_off;   // pairs; capacity L
};

__device__ __forceinline__ int trace_ref_byte(const unsigned char* raw, const TraceJob& jb, int i) // i = 1..R
{
	return raw[jb.ref_off + (jb.reversed ? jb.R - i : i - 1)];
}
__device__ __forceinline__ int trace_read_byte(const unsigned char* raw, const TraceJob& jb, int j) // j = 1..L
{
	return raw[jb.read_off + (jb.reversed ? jb.L - j : j - 1)];
}

__global__ void __launch_bounds__(128) trace_fill_kernel(const unsigned char* __restrict__ raw, const TraceJob* __restrict__ jobs,
                                                         int n_jobs, int match, int mismatch, int gap, int end_gaps,
                                                         unsigned* __restrict__ bits, int* __restrict__ scratch)
{
	const int lane = threadIdx.x & 31;
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int job = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; job < n_jobs; job += warps)
	{
		const TraceJob jb = jobs[job];
		const int R = jb.R, L = jb.L;
		const int W = (R + 16) >> 4;
		unsigned* const my_bits = bits + jb.bits_off;
		int* const edge = scratch + jb.scratch_off;
		const int row0 = end_gaps ? 0 : gap; // H(0,j) = j * row0 (:44-48)
		for (int tile = 0; tile < L; tile += 32)
		{
			const int j = tile + lane + 1;
			const bool live = j <= L;
			const int rc = live ? trace_read_byte(raw, jb, j) : -1;
			int own = j * row0;        // H(i-1, j), starts at column 0
			int diag = (j - 1) * row0; // H(i-1, j-1)
			int cur = 0;               // H(i, j) just computed: what the lane below sees as H(i, j-1)
			unsigned word = 0;
			unsigned* const row_bits = my_bits + (long long)j * W;
			const bool last_lane = lane == 31 && tile + 32 < L;
			for (int s = 1; s <= R + 31; s++)
			{
				const int i = s - lane;
				int above = __shfl_up_sync(0xffffffffu, cur, 1);
				if (live && i >= 1 && i <= R)
				{
					if (lane == 0) above = tile == 0 ? 0 : edge[i]; // H(i, tile): row 0 is all zero (:40-43)
					const int m = diag + (trace_ref_byte(raw, jb, i) == rc ? match : mismatch);
					const int gr = own + gap;   // from (i-1, j)
					const int gd = above + gap; // from (i, j-1)
					const int mx = max(m, max(gr, gd));
					const unsigned dir = gd == mx ? 2u : (gr == mx ? 1u : 0u);
					word |= dir << (2 * (i & 15));
					if ((i & 15) == 15 || i == R)
					{
						row_bits[i >> 4] = word;
						word = 0;
					}
					diag = above;
					own = mx;
					cur = mx;
					if (last_lane) edge[i] = mx;
				}
			}
			__syncwarp();
		}
	}
}

__global__ void trace_walk_kernel(const TraceJob* __restrict__ jobs, int n_jobs, const unsigned* __restrict__ bits,
                                  int2* __restrict__ matches, int* __restrict__ n_matches)
{
	const int job = blockIdx.x * blockDim.x + threadIdx.x;
	if (job >= n_jobs) return;
	const TraceJob jb = jobs[job];
	const int W = (jb.R + 16) >> 4;
	const unsigned* const my_bits = bits + jb.bits_off;
	int2* out = matches + jb.match_off;
	int i = jb.start_i, j = jb.start_j, n = 0;
	while (j > 0)
	{
		// row i == 0 points along the read (:46-47)
		const unsigned dir = i == 0 ? 2u : (my_bits[(long long)j * W + (i >> 4)] >> (2 * (i & 15))) & 3u;
		if (dir == 0)
		{
			out[n++] = make_int2(i - 1, j - 1); // (refPos, readPos) of a diagonal step (:133-139)
			i--;
			j--;
		}
		else if (dir == 1)
			i--;
		else
			j--;
	}
	n_matches[job] = n;
}
}  // namespace dfb

// ------------------------------------------------------------------------------------------
// host entry
// ------------------------------------------------------------------------------------------

extern "C" int dfb_split_backtrace_batch(dfb_ctx* ctx, const dfb_split_params* params, const dfb_seq_table* refs,
                                         const dfb_seq_table* reads, const int32_t* task_cluster, const int32_t* task_read,
                                         const int32_t* task_ref_split1, const int32_t* task_ref_split2,
                                         const int32_t* task_read_split, int64_t n_tasks, int64_t* match_off,
                                         int32_t* matches, int64_t matches_cap, int64_t* n_pairs_out)
{
	if (!ctx) return DFB_ERR_ARG;
	if (!params || n_tasks < 0 || !match_off || !n_pairs_out ||
	    (n_tasks > 0 && (!task_cluster || !task_read || !task_ref_split1 || !task_ref_split2 || !task_read_split)))
		return set_err(ctx, DFB_ERR_ARG, "dfb_split_backtrace_batch: null argument");
	int rc = check_table(ctx, refs, "refs");
	if (rc) return rc;
	rc = check_table(ctx, reads, "reads");
	if (rc) return rc;
	if (refs->n % 2) return set_err(ctx, DFB_ERR_ARG, "refs must hold two windows per cluster");
	CK(ctx, cudaSetDevice(ctx->device));

	// validate and size: the start cells of task t are (i1, a) in matrix 1 and (R2 - refSplit2 - 1, L - a) in matrix 2
	match_off[0] = 0;
	for (int64_t t = 0; t < n_tasks; t++)
	{
		const int64_t c = task_cluster[t], r = task_read[t];
		if (c < 0 || 2 * c + 1 >= refs->n || r < 0 || r >= reads->n)
			return set_err(ctx, DFB_ERR_ARG, "task %lld: index out of range", (long long)t);
		const int64_t R1 = refs->off[2 * c + 1] - refs->off[2 * c], R2 = refs->off[2 * c + 2] - refs->off[2 * c + 1];
		const int64_t L = reads->off[r + 1] - reads->off[r];
		const int64_t a = task_read_split[t], i1 = task_ref_split1[t], i2 = R2 - (int64_t)task_ref_split2[t] - 1;
		if (a < 0 || a > L || i1 < 0 || i1 > R1 || i2 < 0 || i2 > R2)
			return set_err(ctx, DFB_ERR_ARG, "task %lld: start cell outside the matrices", (long long)t);
		match_off[2 * t + 1] = match_off[2 * t] + a;       // a diagonal step consumes one read base
		match_off[2 * t + 2] = match_off[2 * t + 1] + (L - a);
	}
	// (match_off holds capacities for now; it is compacted to the real counts below)
	const int64_t cap_pairs = match_off[2 * n_tasks];
	std::vector<int64_t> cap_off(match_off, match_off + 2 * n_tasks + 1);

	// chunks bounded by pointer-matrix memory
	const size_t kMaxBitsWords = (size_t)1 << 28; // 1 GiB of pointer words per chunk
	std::vector<TraceJob> jobs;
	std::vector<int2> h_pairs;
	std::vector<int> h_cnt;
	std::vector<int64_t> counts((size_t)(2 * n_tasks), 0);
	std::vector<int2> all_pairs((size_t)cap_pairs);
	const int64_t raw_ref = refs->off[refs->n], raw_read = reads->off[reads->n];
	unsigned char* d_raw = nullptr;
	cudaStream_t st = ctx->stream;
	CK(ctx, cudaMallocAsync((void**)&d_raw, (size_t)(raw_ref + raw_read + 16), st));
	auto fail = [&](int code) {
		cudaStreamSynchronize(st);
		cudaFreeAsync(d_raw, st);
		return code;
	};
	cudaError_t e = cudaSuccess;
	if (raw_ref) e = cudaMemcpyAsync(d_raw, refs->bytes, (size_t)raw_ref, cudaMemcpyHostToDevice, st);
	if (e == cudaSuccess && raw_read) e = cudaMemcpyAsync(d_raw + raw_ref, reads->bytes, (size_t)raw_read, cudaMemcpyHostToDevice, st);
	if (e != cudaSuccess) return fail(set_err(ctx, DFB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)));

	for (int64_t first = 0; first < n_tasks;)
	{
		jobs.clear();
		size_t words = 0, scratch = 0, pairs = 0;
		int64_t last = first;
		for (; last < n_tasks; last++)
		{
			const int64_t c = task_cluster[last], r = task_read[last];
			const int64_t R1 = refs->off[2 * c + 1] - refs->off[2 * c], R2 = refs->off[2 * c + 2] - refs->off[2 * c + 1];
			const int64_t L = reads->off[r + 1] - reads->off[r];
			const size_t w1 = (size_t)(L + 1) * (size_t)((R1 + 16) >> 4), w2 = (size_t)(L + 1) * (size_t)((R2 + 16) >> 4);
			if (!jobs.empty() && words + w1 + w2 > kMaxBitsWords) break;
			const int a = task_read_split[last];
			TraceJob j1{refs->off[2 * c], raw_ref + reads->off[r], (int)R1, (int)L, 0, task_ref_split1[last], a,
			            (long long)words, (long long)scratch, (long long)pairs};
			words += w1;
			scratch += (size_t)R1 + 1;
			pairs += (size_t)a;
			TraceJob j2{refs->off[2 * c + 1], raw_ref + reads->off[r], (int)R2, (int)L, 1,
			            (int)(R2 - task_ref_split2[last] - 1), (int)(L - a), (long long)words, (long long)scratch, (long long)pairs};
			words += w2;
			scratch += (size_t)R2 + 1;
			pairs += (size_t)(L - a);
			jobs.push_back(j1);
			jobs.push_back(j2);
		}
		const int n_jobs = (int)jobs.size();
		TraceJob* d_jobs = nullptr;
		unsigned* d_bits = nullptr;
		int* d_scratch = nullptr;
		int2* d_pairs = nullptr;
		int* d_cnt = nullptr;
		if ((e = cudaMallocAsync((void**)&d_jobs, sizeof(TraceJob) * (size_t)n_jobs, st)) != cudaSuccess ||
		    (e = cudaMallocAsync((void**)&d_bits, sizeof(unsigned) * (words + 1), st)) != cudaSuccess ||
		    (e = cudaMallocAsync((void**)&d_scratch, sizeof(int) * (scratch + 1), st)) != cudaSuccess ||
		    (e = cudaMallocAsync((void**)&d_pairs, sizeof(int2) * (pairs + 1), st)) != cudaSuccess ||
		    (e = cudaMallocAsync((void**)&d_cnt, sizeof(int) * (size_t)n_jobs, st)) != cudaSuccess)
		{
			// (blocks already taken go back to the pool with the stream)
			if (d_jobs) cudaFreeAsync(d_jobs, st);
			if (d_bits) cudaFreeAsync(d_bits, st);
			if (d_scratch) cudaFreeAsync(d_scratch, st);
			if (d_pairs) cudaFreeAsync(d_pairs, st);
			return fail(set_err(ctx, DFB_ERR_NOMEM, "device allocation failed: %s", cudaGetErrorString(e)));
		}
		h_pairs.resize(pairs + 1);
		h_cnt.resize((size_t)n_jobs);
		e = cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(TraceJob) * (size_t)n_jobs, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess)
		{
			const int warps_per_block = 4;
			const int grid = std::max(1, std::min((n_jobs + warps_per_block - 1) / warps_per_block, ctx->prop.multiProcessorCount * 8));
			trace_fill_kernel<<<grid, 32 * warps_per_block, 0, st>>>(d_raw, d_jobs, n_jobs, params->match, params->mismatch,
			                                                         params->gap, params->end_gaps, d_bits, d_scratch);
			trace_walk_kernel<<<(n_jobs + 63) / 64, 64, 0, st>>>(d_jobs, n_jobs, d_bits, d_pairs, d_cnt);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaMemcpyAsync(h_pairs.data(), d_pairs, sizeof(int2) * pairs, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt.data(), d_cnt, sizeof(int) * (size_t)n_jobs, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		cudaFreeAsync(d_jobs, st);
		cudaFreeAsync(d_bits, st);
		cudaFreeAsync(d_scratch, st);
		cudaFreeAsync(d_pairs, st);
		cudaFreeAsync(d_cnt, st);
		if (e != cudaSuccess) return fail(set_err(ctx, DFB_ERR_CUDA, "backtrace kernels failed: %s", cudaGetErrorString(e)));
		for (int k = 0; k < n_jobs; k++)
		{
			const int64_t t = first + k / 2;
			const int half = k & 1;
			const TraceJob& jb = jobs[(size_t)k];
			const int n = h_cnt[(size_t)k];
			counts[(size_t)(2 * t + half)] = n;
			int2* dst = all_pairs.data() + cap_off[(size_t)(2 * t + half)];
			const int2* src = h_pairs.data() + jb.match_off;
			if (!half)
				for (int q = 0; q < n; q++) dst[q] = src[n - 1 - q]; // BackTracePath's reverse (:142)
			else
				for (int q = 0; q < n; q++) dst[q] = make_int2(jb.R - src[q].x - 1, jb.L - src[q].y - 1); // ReverseMatches (:145-154)
		}
		first = last;
	}
	cudaFreeAsync(d_raw, st);

	// compact: real counts instead of capacities
	int64_t total = 0;
	for (int64_t k = 0; k < 2 * n_tasks; k++)
	{
		match_off[k] = total;
		total += counts[(size_t)k];
	}
	match_off[2 * n_tasks] = total;
	*n_pairs_out = total;
	if (matches)
	{
		if (matches_cap < total) return set_err(ctx, DFB_ERR_ARG, "matches buffer holds %lld pairs, %lld needed", (long long)matches_cap, (long long)total);
		for (int64_t k = 0; k < 2 * n_tasks; k++)
		{
			const int2* src = all_pairs.data() + cap_off[(size_t)k];
			int32_t* dst = matches + 2 * match_off[k];
			for (int64_t q = 0; q < counts[(size_t)k]; q++)
			{
				dst[2 * q] = src[q].x;
				dst[2 * q + 1] = src[q].y;
			}
		}
	}
	return DFB_OK;
}
