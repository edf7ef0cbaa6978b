// Host side of the C ABI in include/defuse_b200.h: context, plan building (packing,
// length-bucketed job lists), kernel dispatch and result assembly.  No CPU implementation
// of the DP exists in this library: without a GPU every entry point fails.
#include "../../include/defuse_b200.h"
#include "dfb_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace dfb;

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------

struct dfb_ctx
{
	int device = 0;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	cudaDeviceProp prop;
	mutable std::string err;
	dfb_plan* last_split = nullptr; // result holder of dfb_split_align_batch
};

static thread_local std::string g_create_err;

static int set_err(const dfb_ctx* ctx, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) ctx->err = buf; else g_create_err = buf;
	return code;
}

#define CK(ctx, call)                                                                                  \
	do                                                                                                 \
	{                                                                                                  \
		cudaError_t e__ = (call);                                                                      \
		if (e__ != cudaSuccess)                                                                        \
			return set_err(ctx, DFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
			               __FILE__, __LINE__);                                                        \
	} while (0)

extern "C" int dfb_abi_version(void) { return DFB_ABI_VERSION; }

extern "C" int dfb_device_count(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess)
	{
		set_err(nullptr, DFB_ERR_NODEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
		return -DFB_ERR_NODEVICE;
	}
	return n;
}

extern "C" int dfb_ctx_create(int device_ordinal, dfb_ctx** out)
{
	if (!out) return set_err(nullptr, DFB_ERR_ARG, "dfb_ctx_create: null output pointer");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
	{
		return set_err(nullptr, DFB_ERR_NODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
		               e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	}
	if (device_ordinal < 0 || device_ordinal >= n)
	{
		return set_err(nullptr, DFB_ERR_NODEVICE, "device ordinal %d out of range (0..%d)", device_ordinal, n - 1);
	}
	dfb_ctx* ctx = new (std::nothrow) dfb_ctx();
	if (!ctx) return set_err(nullptr, DFB_ERR_NOMEM, "out of host memory");
	ctx->device = device_ordinal;
	e = cudaGetDeviceProperties(&ctx->prop, device_ordinal);
	if (e != cudaSuccess)
	{
		delete ctx;
		return set_err(nullptr, DFB_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
	}
	if (ctx->prop.major != 10)
	{
		int major = ctx->prop.major, minor = ctx->prop.minor;
		delete ctx;
		return set_err(nullptr, DFB_ERR_NODEVICE,
		               "device %d is compute capability %d.%d; the kernels are built for sm_100a only", device_ordinal,
		               major, minor);
	}
	if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess ||
	    (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
	{
		delete ctx;
		return set_err(nullptr, DFB_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(e));
	}
	ctx->stream = ctx->own_stream;
	*out = ctx;
	return DFB_OK;
}

extern "C" void dfb_ctx_destroy(dfb_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->last_split) dfb_plan_destroy(ctx->last_split);
	if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

extern "C" const char* dfb_last_error(const dfb_ctx* ctx)
{
	return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int dfb_ctx_set_stream(dfb_ctx* ctx, void* cuda_stream)
{
	if (!ctx) return DFB_ERR_ARG;
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
	return DFB_OK;
}

extern "C" int dfb_ctx_device_info(const dfb_ctx* ctx, dfb_device_info* info)
{
	if (!ctx || !info) return DFB_ERR_ARG;
	memset(info, 0, sizeof(*info));
	snprintf(info->name, sizeof(info->name), "%s", ctx->prop.name);
	info->ordinal = ctx->device;
	info->sm_count = ctx->prop.multiProcessorCount;
	info->cc_major = ctx->prop.major;
	info->cc_minor = ctx->prop.minor;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
	info->clock_khz = khz;
	info->total_mem = (int64_t)ctx->prop.totalGlobalMem;
	return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// kernel classes: (lanes per job pair, rows per lane); capacity G*S read rows
// ------------------------------------------------------------------------------------------

#define DFB_CLASSES(X) \
	X(8, 4) X(8, 7) X(8, 10) X(8, 13) X(8, 16) X(8, 19) X(8, 22) X(8, 25) X(8, 32) \
	X(16, 20) X(16, 25) X(16, 32) X(32, 24) X(32, 32)

struct ClassDef
{
	int G, S;
};
#define X(g, s) {g, s},
static const ClassDef kClasses[] = {DFB_CLASSES(X)};
#undef X
static const int kNumClasses = (int)(sizeof(kClasses) / sizeof(kClasses[0]));
static const int kMaxFastRows = 1024;

template <int G, int S, int MODE>
static cudaError_t launch_fast_t(const FastParams& p, int sm_count, cudaStream_t stream, int n_items_bound)
{
	static int occ = 0;
	if (occ == 0)
	{
		int o = 0;
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dp_fast_kernel<G, S, MODE>, 128, 0);
		if (e != cudaSuccess) return e;
		occ = o > 0 ? o : 1;
	}
	const int per_block = 4 * (32 / G);
	long long want = ((long long)n_items_bound + per_block - 1) / per_block;
	long long cap = (long long)sm_count * occ;
	int grid = (int)std::max(1LL, std::min(want, cap));
	dp_fast_kernel<G, S, MODE><<<grid, 128, 0, stream>>>(p);
	return cudaGetLastError();
}

static cudaError_t launch_fast(int cls, int mode, const FastParams& p, int sm_count, cudaStream_t stream, int bound)
{
	int idx = 0;
#define X(g, s)                                                                                   \
	if (idx++ == cls)                                                                             \
	{                                                                                             \
		if (mode == MODE_SIMPLE) return launch_fast_t<g, s, MODE_SIMPLE>(p, sm_count, stream, bound); \
		if (mode == MODE_SPLIT) return launch_fast_t<g, s, MODE_SPLIT>(p, sm_count, stream, bound);   \
		return launch_fast_t<g, s, MODE_PROBE>(p, sm_count, stream, bound);                           \
	}
	DFB_CLASSES(X)
#undef X
	return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------

struct ClassWork
{
	std::vector<JobPair> jobs;
	JobPair* d_jobs = nullptr;
	int* d_ctrl = nullptr; // [0] cursor, [1] hit_count, [2] probe cursor
	int* d_hitq = nullptr;
	uint32_t* d_ntg = nullptr;
	FastParams fp;
};

struct dfb_plan
{
	dfb_ctx* ctx = nullptr;
	bool split = false;
	dfb_split_params sp{};
	int64_t n_tasks = 0;
	bool fast_ok[kNumClasses];
	uint32_t bias[kNumClasses];

	// device
	uint8_t* d_raw = nullptr;
	PackItem* d_items = nullptr;
	uint2* d_pool = nullptr;
	uint8_t* d_obytes = nullptr;
	int32_t* d_out = nullptr; // score / best per task
	Event* d_events = nullptr;
	unsigned long long* d_ev_count = nullptr;
	unsigned long long ev_cap = 0;
	ClassWork cls[kNumClasses];
	// generic path
	std::vector<GenJob> gen_jobs;
	GenJob* d_gen_jobs = nullptr;
	int* d_gen_ctrl = nullptr; // [0] cursor pass 1, [1] cursor probe
	int32_t* d_gen_rowmax = nullptr;
	uint8_t* d_gen_row_en = nullptr;
	int32_t* d_gen_bnd = nullptr;
	int64_t gen_bnd_stride = 0;
	int gen_grid = 0;
	int* d_gen_probe_flag = nullptr;
	int32_t* d_task_min_score = nullptr;
	int64_t gen_rows_total = 0;

	// host
	std::vector<int32_t> task_L;      // read length per task (split)
	std::vector<dfb_split_row> rows;
	std::vector<int32_t> cols;
	bool ran = false, fetched = false;
	bool timing = false;
	bool pack_timed = false;
	bool run_timed = false;
	cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
	dfb_plan_stats stats{};
};

template <class T>
static void dfree(T*& p)
{
	if (p) cudaFree(p);
	p = nullptr;
}

extern "C" void dfb_plan_destroy(dfb_plan* plan)
{
	if (!plan) return;
	if (plan->ctx)
	{
		cudaSetDevice(plan->ctx->device);
		if (plan->ctx->last_split == plan) plan->ctx->last_split = nullptr;
	}
	dfree(plan->d_raw);
	dfree(plan->d_items);
	dfree(plan->d_pool);
	dfree(plan->d_obytes);
	dfree(plan->d_out);
	dfree(plan->d_events);
	dfree(plan->d_ev_count);
	for (int c = 0; c < kNumClasses; c++)
	{
		dfree(plan->cls[c].d_jobs);
		dfree(plan->cls[c].d_ctrl);
		dfree(plan->cls[c].d_hitq);
		dfree(plan->cls[c].d_ntg);
	}
	dfree(plan->d_gen_jobs);
	dfree(plan->d_gen_ctrl);
	dfree(plan->d_gen_rowmax);
	dfree(plan->d_gen_row_en);
	dfree(plan->d_gen_bnd);
	dfree(plan->d_gen_probe_flag);
	dfree(plan->d_task_min_score);
	for (int k = 0; k < 3; k++)
		if (plan->ev[k]) cudaEventDestroy(plan->ev[k]);
	delete plan;
}

static int check_table(const dfb_ctx* ctx, const dfb_seq_table* t, const char* what)
{
	if (!t || !t->off || t->n < 0 || (t->n > 0 && !t->bytes && t->off[t->n] > 0))
		return set_err(ctx, DFB_ERR_ARG, "%s: null table", what);
	if (t->off[0] != 0) return set_err(ctx, DFB_ERR_ARG, "%s: off[0] must be 0", what);
	for (int64_t k = 0; k < t->n; k++)
	{
		if (t->off[k + 1] < t->off[k]) return set_err(ctx, DFB_ERR_ARG, "%s: offsets decrease at %lld", what, (long long)k);
		if (t->off[k + 1] - t->off[k] > 0x7fffffffLL) return set_err(ctx, DFB_ERR_ARG, "%s: sequence %lld too long", what, (long long)k);
	}
	return DFB_OK;
}

// Scoring triples the s16x2 kernels are exact for (DESIGN.md "number range"):
//   match >= 1, mismatch <= 0, gap <= 0  =>  j*gap <= H(i,j) <= j*match  for every cell,
// so H - match*j fits 16 bits for the class' row capacity.  Everything else goes to s32.
static void classify_params(dfb_plan* pl, int m, int x, int g, bool extra_ok)
{
	for (int c = 0; c < kNumClasses; c++)
	{
		const int rows = kClasses[c].G * kClasses[c].S;
		bool ok = extra_ok && m >= 1 && x <= 0 && g <= 0;
		long long range = (long long)rows * ((long long)m - g) + ((long long)m - x);
		long long B = range + 8;
		if (B > 32000 || B + (long long)m * kClasses[c].S > 32767 || (long long)m * rows > 32767) ok = false;
		pl->fast_ok[c] = ok;
		pl->bias[c] = ok ? (uint32_t)B : 0;
	}
}

static int class_for_rows(int L)
{
	for (int c = 0; c < kNumClasses; c++)
		if (kClasses[c].G * kClasses[c].S >= L) return c;
	return -1;
}

static uint32_t pack2(int v) { return ((uint32_t)v & 0xFFFFu) | ((uint32_t)v << 16); }

struct SeqLayout
{
	std::vector<PackItem> items;
	uint32_t total_words = 0;
	// adds one stored sequence, returns its first pool word
	uint32_t add(int64_t src, uint32_t len, uint32_t flags)
	{
		PackItem it;
		it.src = src;
		it.len = len;
		it.dst_word = total_words;
		it.flags = flags;
		it.pad = 0;
		items.push_back(it);
		uint32_t w = total_words;
		total_words += (len + 15) / 16;
		return w;
	}
};

static int upload_and_pack(dfb_plan* pl, const dfb_seq_table* a, const dfb_seq_table* b, SeqLayout& lay)
{
	dfb_ctx* ctx = pl->ctx;
	const int64_t na = a->off[a->n], nb = b->off[b->n];
	CK(ctx, cudaMalloc(&pl->d_raw, (size_t)std::max<int64_t>(na + nb, 16)));
	if (na) CK(ctx, cudaMemcpyAsync(pl->d_raw, a->bytes, (size_t)na, cudaMemcpyHostToDevice, ctx->stream));
	if (nb) CK(ctx, cudaMemcpyAsync(pl->d_raw + na, b->bytes, (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
	const size_t n_items = lay.items.size();
	// one spare word so that an empty batch still has valid pointers
	const uint32_t words = lay.total_words + 1;
	CK(ctx, cudaMalloc(&pl->d_items, std::max<size_t>(n_items, 1) * sizeof(PackItem)));
	CK(ctx, cudaMalloc(&pl->d_pool, (size_t)words * sizeof(uint2)));
	CK(ctx, cudaMalloc(&pl->d_obytes, (size_t)words * 16));
	if (n_items)
		CK(ctx, cudaMemcpyAsync(pl->d_items, lay.items.data(), n_items * sizeof(PackItem), cudaMemcpyHostToDevice, ctx->stream));
	if (lay.total_words && n_items)
	{
		int grid = (int)std::min<uint32_t>((lay.total_words + 255) / 256, 148 * 16);
		for (int k = 0; k < 3; k++) CK(ctx, cudaEventCreate(&pl->ev[k]));
		CK(ctx, cudaEventRecord(pl->ev[0], ctx->stream));
		pack_kernel<<<grid, 256, 0, ctx->stream>>>(pl->d_raw, pl->d_items, (int)n_items, lay.total_words, pl->d_pool, pl->d_obytes);
		CK(ctx, cudaGetLastError());
		CK(ctx, cudaEventRecord(pl->ev[1], ctx->stream));
		pl->pack_timed = true;
	}
	pl->stats.h2d_bytes += na + nb + (int64_t)(n_items * sizeof(PackItem));
	pl->stats.raw_bytes = na + nb;
	pl->stats.packed_bytes = (int64_t)lay.total_words * 8;
	return DFB_OK;
}

static int upload_jobs(dfb_plan* pl, bool split)
{
	dfb_ctx* ctx = pl->ctx;
	for (int c = 0; c < kNumClasses; c++)
	{
		ClassWork& cw = pl->cls[c];
		if (cw.jobs.empty()) continue;
		const size_t n = cw.jobs.size();
		CK(ctx, cudaMalloc(&cw.d_jobs, n * sizeof(JobPair)));
		CK(ctx, cudaMemcpyAsync(cw.d_jobs, cw.jobs.data(), n * sizeof(JobPair), cudaMemcpyHostToDevice, ctx->stream));
		CK(ctx, cudaMalloc(&cw.d_ctrl, 4 * sizeof(int)));
		pl->stats.h2d_bytes += (int64_t)(n * sizeof(JobPair));
		pl->stats.fast_jobs += (int64_t)n;
		if (split)
		{
			CK(ctx, cudaMalloc(&cw.d_hitq, n * sizeof(int)));
			CK(ctx, cudaMalloc(&cw.d_ntg, n * (size_t)kClasses[c].G * kClasses[c].S * sizeof(uint32_t)));
		}
	}
	if (!pl->gen_jobs.empty())
	{
		const size_t n = pl->gen_jobs.size();
		CK(ctx, cudaMalloc(&pl->d_gen_jobs, n * sizeof(GenJob)));
		CK(ctx, cudaMemcpyAsync(pl->d_gen_jobs, pl->gen_jobs.data(), n * sizeof(GenJob), cudaMemcpyHostToDevice, ctx->stream));
		CK(ctx, cudaMalloc(&pl->d_gen_ctrl, 4 * sizeof(int)));
		uint32_t maxR = 0;
		for (const GenJob& j : pl->gen_jobs) maxR = std::max(maxR, j.R);
		pl->gen_bnd_stride = (int64_t)maxR + 2;
		pl->gen_grid = (int)std::min<size_t>((n + 3) / 4, (size_t)ctx->prop.multiProcessorCount * 4);
		CK(ctx, cudaMalloc(&pl->d_gen_bnd, (size_t)pl->gen_grid * 4 * 2 * pl->gen_bnd_stride * sizeof(int32_t)));
		if (split)
		{
			CK(ctx, cudaMalloc(&pl->d_gen_rowmax, (size_t)std::max<int64_t>(pl->gen_rows_total, 1) * sizeof(int32_t)));
			CK(ctx, cudaMalloc(&pl->d_gen_row_en, (size_t)std::max<int64_t>(pl->gen_rows_total, 1)));
			CK(ctx, cudaMemsetAsync(pl->d_gen_row_en, 0, (size_t)std::max<int64_t>(pl->gen_rows_total, 1), ctx->stream));
			CK(ctx, cudaMalloc(&pl->d_gen_probe_flag, (n / 2 + 1) * sizeof(int)));
		}
		pl->stats.h2d_bytes += (int64_t)(n * sizeof(GenJob));
		pl->stats.generic_jobs += (int64_t)n;
	}
	return DFB_OK;
}

static void fill_fast_params(dfb_plan* pl, int c, int m, int x, int g, int min_split)
{
	ClassWork& cw = pl->cls[c];
	FastParams& fp = cw.fp;
	memset(&fp, 0, sizeof(fp));
	fp.pool = pl->d_pool;
	fp.obytes = pl->d_obytes;
	fp.jobs = cw.d_jobs;
	fp.n_jobs = (int)cw.jobs.size();
	fp.m = m;
	fp.bias = pl->bias[c];
	fp.xm = (uint32_t)(x - m);
	fp.g2 = pack2(g);
	fp.gm2 = pack2(g - m);
	fp.min_split = min_split;
	fp.out = pl->d_out;
	fp.hit_count = cw.d_ctrl ? cw.d_ctrl + 1 : nullptr;
	fp.hitq = cw.d_hitq;
	fp.ntg = cw.d_ntg;
	fp.events = pl->d_events;
	fp.ev_count = pl->d_ev_count;
	fp.ev_cap = pl->ev_cap;
	for (int k = 0; k < 32; k++) fp.ck[k] = pack2(m * (k + 1));
}

// ---- SimpleAligner plan ------------------------------------------------------------------

extern "C" int dfb_simple_plan_create(dfb_ctx* ctx, const dfb_simple_params* params, const dfb_seq_table* refs,
                                      const dfb_seq_table* seqs, const int32_t* task_ref, const int32_t* task_seq,
                                      int64_t n_tasks, dfb_plan** out)
{
	if (!ctx) return DFB_ERR_ARG;
	if (!params || !out || n_tasks < 0 || (n_tasks > 0 && (!task_ref || !task_seq)))
		return set_err(ctx, DFB_ERR_ARG, "dfb_simple_plan_create: null argument");
	*out = nullptr;
	int rc;
	if ((rc = check_table(ctx, refs, "refs")) || (rc = check_table(ctx, seqs, "seqs"))) return rc;
	if (n_tasks > 0x7fffffffLL) return set_err(ctx, DFB_ERR_ARG, "too many tasks in one batch");
	for (int64_t t = 0; t < n_tasks; t++)
	{
		if (task_ref[t] < 0 || task_ref[t] >= refs->n || task_seq[t] < 0 || task_seq[t] >= seqs->n)
			return set_err(ctx, DFB_ERR_ARG, "task %lld: table index out of range", (long long)t);
	}
	CK(ctx, cudaSetDevice(ctx->device));
	dfb_plan* pl = new (std::nothrow) dfb_plan();
	if (!pl) return set_err(ctx, DFB_ERR_NOMEM, "out of host memory");
	pl->ctx = ctx;
	pl->split = false;
	pl->n_tasks = n_tasks;
	pl->stats.n_tasks = n_tasks;
	classify_params(pl, params->match, params->mismatch, params->gap, true);

	SeqLayout lay;
	std::vector<uint32_t> ref_word((size_t)refs->n), seq_word((size_t)seqs->n);
	for (int64_t k = 0; k < refs->n; k++) ref_word[k] = lay.add(refs->off[k], (uint32_t)(refs->off[k + 1] - refs->off[k]), 0);
	const int64_t seq_base = refs->off[refs->n];
	for (int64_t k = 0; k < seqs->n; k++) seq_word[k] = lay.add(seq_base + seqs->off[k], (uint32_t)(seqs->off[k + 1] - seqs->off[k]), 0);

	// bucket by kernel class, longest reference first inside a bucket, pair neighbours
	struct Key
	{
		uint64_t key;
		int32_t task;
	};
	std::vector<Key> keys;
	keys.reserve((size_t)n_tasks);
	for (int64_t t = 0; t < n_tasks; t++)
	{
		const int64_t R = refs->off[task_ref[t] + 1] - refs->off[task_ref[t]];
		const int64_t L = seqs->off[task_seq[t] + 1] - seqs->off[task_seq[t]];
		pl->stats.cells += R * L;
		if (R == 0 || L == 0) continue; // no interior cell: score 0 (d_out is zero-initialised)
		int c = (L <= kMaxFastRows && R <= 65535) ? class_for_rows((int)L) : -1;
		if (c >= 0 && !pl->fast_ok[c]) c = -1;
		if (c < 0)
		{
			GenJob j;
			j.ref_w = ref_word[task_ref[t]];
			j.read_w = seq_word[task_seq[t]];
			j.R = (uint32_t)R;
			j.L = (uint32_t)L;
			j.task = (int32_t)t;
			j.half = 0;
			j.row_off = 0;
			pl->gen_jobs.push_back(j);
			continue;
		}
		Key k;
		k.key = ((uint64_t)c << 48) | ((uint64_t)(65535 - R) << 24) | (uint64_t)(0xFFFFFF - L);
		k.task = (int32_t)t;
		keys.push_back(k);
	}
	std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.key < b.key || (a.key == b.key && a.task < b.task); });
	for (size_t i = 0; i < keys.size();)
	{
		const int c = (int)(keys[i].key >> 48);
		JobPair jp;
		memset(&jp, 0, sizeof(jp));
		jp.out0 = jp.out1 = -1;
		for (int h = 0; h < 2 && i < keys.size() && (int)(keys[i].key >> 48) == c; h++, i++)
		{
			const int32_t t = keys[i].task;
			jp.ref_w[h] = ref_word[task_ref[t]];
			jp.read_w[h] = seq_word[task_seq[t]];
			jp.R[h] = (uint16_t)(refs->off[task_ref[t] + 1] - refs->off[task_ref[t]]);
			jp.L[h] = (uint16_t)(seqs->off[task_seq[t] + 1] - seqs->off[task_seq[t]]);
			if (h == 0) jp.out0 = t; else jp.out1 = t;
		}
		pl->cls[c].jobs.push_back(jp);
	}

	rc = upload_and_pack(pl, refs, seqs, lay);
	if (!rc) rc = upload_jobs(pl, false);
	if (!rc)
	{
		cudaError_t e = cudaMalloc(&pl->d_out, (size_t)std::max<int64_t>(n_tasks, 1) * sizeof(int32_t));
		if (e == cudaSuccess) e = cudaMemsetAsync(pl->d_out, 0, (size_t)std::max<int64_t>(n_tasks, 1) * sizeof(int32_t), ctx->stream);
		if (e != cudaSuccess) rc = set_err(ctx, DFB_ERR_CUDA, "output allocation failed: %s", cudaGetErrorString(e));
	}
	if (!rc)
	{
		for (int c = 0; c < kNumClasses; c++)
			if (!pl->cls[c].jobs.empty()) fill_fast_params(pl, c, params->match, params->mismatch, params->gap, 0);
		pl->sp.match = params->match;
		pl->sp.mismatch = params->mismatch;
		pl->sp.gap = params->gap;
		cudaError_t e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) rc = set_err(ctx, DFB_ERR_CUDA, "plan upload/pack failed: %s", cudaGetErrorString(e));
		if (!rc && pl->pack_timed)
		{
			float ms = 0;
			if (cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]) == cudaSuccess) pl->stats.ms_pack = ms;
		}
	}
	if (rc)
	{
		dfb_plan_destroy(pl);
		return rc;
	}
	*out = pl;
	return DFB_OK;
}

// ---- SplitReadAligner plan ------------------------------------------------------------------

extern "C" int dfb_split_plan_create(dfb_ctx* ctx, const dfb_split_params* params, const dfb_seq_table* refs,
                                     const dfb_seq_table* reads, const int32_t* task_cluster, const int32_t* task_read,
                                     const int32_t* task_min_score, int64_t n_tasks, dfb_plan** out)
{
	if (!ctx) return DFB_ERR_ARG;
	if (!params || !out || n_tasks < 0 || (n_tasks > 0 && (!task_cluster || !task_read || !task_min_score)))
		return set_err(ctx, DFB_ERR_ARG, "dfb_split_plan_create: null argument");
	*out = nullptr;
	int rc;
	if ((rc = check_table(ctx, refs, "refs")) || (rc = check_table(ctx, reads, "reads"))) return rc;
	if (refs->n & 1) return set_err(ctx, DFB_ERR_ARG, "refs must hold two windows per cluster (n is odd)");
	if (n_tasks > 0x7fffffffLL) return set_err(ctx, DFB_ERR_ARG, "too many tasks in one batch");
	const int64_t n_clusters = refs->n / 2;
	for (int64_t t = 0; t < n_tasks; t++)
	{
		if (task_cluster[t] < 0 || task_cluster[t] >= n_clusters || task_read[t] < 0 || task_read[t] >= reads->n)
			return set_err(ctx, DFB_ERR_ARG, "task %lld: table index out of range", (long long)t);
	}
	CK(ctx, cudaSetDevice(ctx->device));
	dfb_plan* pl = new (std::nothrow) dfb_plan();
	if (!pl) return set_err(ctx, DFB_ERR_NOMEM, "out of host memory");
	pl->ctx = ctx;
	pl->split = true;
	pl->sp = *params;
	pl->n_tasks = n_tasks;
	pl->stats.n_tasks = n_tasks;
	classify_params(pl, params->match, params->mismatch, params->gap, params->end_gaps == 0 && params->min_split_score >= 1);

	// reference 1 of every cluster forward, reference 2 reversed (SplitReadAligner.cpp:80-81);
	// every read forward and reversed (:84-85)
	SeqLayout lay;
	std::vector<uint32_t> ref_word((size_t)refs->n), read_word((size_t)reads->n * 2);
	for (int64_t k = 0; k < refs->n; k++)
		ref_word[k] = lay.add(refs->off[k], (uint32_t)(refs->off[k + 1] - refs->off[k]), (k & 1) ? PACK_REVERSE : 0);
	const int64_t read_base = refs->off[refs->n];
	for (int64_t k = 0; k < reads->n; k++)
	{
		const uint32_t len = (uint32_t)(reads->off[k + 1] - reads->off[k]);
		read_word[2 * k] = lay.add(read_base + reads->off[k], len, 0);
		read_word[2 * k + 1] = lay.add(read_base + reads->off[k], len, PACK_REVERSE);
	}

	struct Key
	{
		uint64_t key;
		int32_t task;
	};
	std::vector<Key> keys;
	keys.reserve((size_t)n_tasks);
	pl->task_L.resize((size_t)n_tasks);
	std::vector<int32_t> gen_tasks;
	for (int64_t t = 0; t < n_tasks; t++)
	{
		const int64_t c2 = 2 * (int64_t)task_cluster[t];
		const int64_t R1 = refs->off[c2 + 1] - refs->off[c2];
		const int64_t R2 = refs->off[c2 + 2] - refs->off[c2 + 1];
		const int64_t L = reads->off[task_read[t] + 1] - reads->off[task_read[t]];
		pl->task_L[t] = (int32_t)L;
		pl->stats.cells += (R1 + R2) * L;
		if (L == 0) continue; // every row maximum is 0: no split (SplitReadAligner.cpp:224-227)
		int c = (L <= kMaxFastRows && R1 <= 65535 && R2 <= 65535) ? class_for_rows((int)L) : -1;
		if (c >= 0 && !pl->fast_ok[c]) c = -1;
		if (c < 0)
		{
			gen_tasks.push_back((int32_t)t);
			continue;
		}
		Key k;
		k.key = ((uint64_t)c << 48) | ((uint64_t)(65535 - std::max(R1, R2)) << 24);
		k.task = (int32_t)t;
		keys.push_back(k);
	}
	std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.key < b.key || (a.key == b.key && a.task < b.task); });
	for (const Key& k : keys)
	{
		const int32_t t = k.task;
		const int c = (int)(k.key >> 48);
		const int64_t c2 = 2 * (int64_t)task_cluster[t];
		JobPair jp;
		jp.ref_w[0] = ref_word[c2];
		jp.ref_w[1] = ref_word[c2 + 1];
		jp.read_w[0] = read_word[2 * (int64_t)task_read[t]];
		jp.read_w[1] = read_word[2 * (int64_t)task_read[t] + 1];
		jp.R[0] = (uint16_t)(refs->off[c2 + 1] - refs->off[c2]);
		jp.R[1] = (uint16_t)(refs->off[c2 + 2] - refs->off[c2 + 1]);
		jp.L[0] = jp.L[1] = (uint16_t)pl->task_L[t];
		jp.out0 = t;
		jp.out1 = task_min_score[t];
		pl->cls[c].jobs.push_back(jp);
	}
	for (int32_t t : gen_tasks)
	{
		const int64_t c2 = 2 * (int64_t)task_cluster[t];
		for (int h = 0; h < 2; h++)
		{
			GenJob j;
			j.ref_w = ref_word[c2 + h];
			j.read_w = read_word[2 * (int64_t)task_read[t] + h];
			j.R = (uint32_t)(refs->off[c2 + h + 1] - refs->off[c2 + h]);
			j.L = (uint32_t)pl->task_L[t];
			j.task = t;
			j.half = h;
			j.row_off = pl->gen_rows_total;
			pl->gen_rows_total += (int64_t)j.L + 1;
			pl->gen_jobs.push_back(j);
		}
	}

	rc = upload_and_pack(pl, refs, reads, lay);
	if (!rc)
	{
		pl->ev_cap = (unsigned long long)std::max<int64_t>(8 * n_tasks, 1 << 20);
		cudaError_t e = cudaMalloc(&pl->d_events, (size_t)pl->ev_cap * sizeof(Event));
		if (e == cudaSuccess) e = cudaMalloc(&pl->d_ev_count, sizeof(unsigned long long));
		if (e == cudaSuccess) e = cudaMalloc(&pl->d_out, (size_t)std::max<int64_t>(n_tasks, 1) * sizeof(int32_t));
		if (e == cudaSuccess) e = cudaMemsetAsync(pl->d_out, 0, (size_t)std::max<int64_t>(n_tasks, 1) * sizeof(int32_t), ctx->stream);
		if (e == cudaSuccess && !pl->gen_jobs.empty())
		{
			e = cudaMalloc(&pl->d_task_min_score, (size_t)n_tasks * sizeof(int32_t));
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(pl->d_task_min_score, task_min_score, (size_t)n_tasks * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
		}
		if (e != cudaSuccess) rc = set_err(ctx, DFB_ERR_CUDA, "output allocation failed: %s", cudaGetErrorString(e));
	}
	if (!rc) rc = upload_jobs(pl, true);
	if (!rc)
	{
		for (int c = 0; c < kNumClasses; c++)
			if (!pl->cls[c].jobs.empty())
				fill_fast_params(pl, c, params->match, params->mismatch, params->gap, params->min_split_score);
		cudaError_t e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) rc = set_err(ctx, DFB_ERR_CUDA, "plan upload/pack failed: %s", cudaGetErrorString(e));
		if (!rc && pl->pack_timed)
		{
			float ms = 0;
			if (cudaEventElapsedTime(&ms, pl->ev[0], pl->ev[1]) == cudaSuccess) pl->stats.ms_pack = ms;
		}
	}
	if (rc)
	{
		dfb_plan_destroy(pl);
		return rc;
	}
	*out = pl;
	return DFB_OK;
}

// ---- run / fetch -------------------------------------------------------------------------------

static GenParams gen_params(dfb_plan* pl, int cursor_slot)
{
	GenParams gp;
	memset(&gp, 0, sizeof(gp));
	gp.obytes = pl->d_obytes;
	gp.jobs = pl->d_gen_jobs;
	gp.n_jobs = (int)pl->gen_jobs.size();
	gp.cursor = pl->d_gen_ctrl + cursor_slot;
	gp.m = pl->sp.match;
	gp.x = pl->sp.mismatch;
	gp.g = pl->sp.gap;
	gp.end_gaps = pl->split ? pl->sp.end_gaps : 0;
	gp.out = pl->d_out;
	gp.rowmax = pl->d_gen_rowmax;
	gp.row_en = pl->d_gen_row_en;
	gp.probe_flag = pl->d_gen_probe_flag;
	gp.bnd = pl->d_gen_bnd;
	gp.bnd_stride = pl->gen_bnd_stride;
	gp.events = pl->d_events;
	gp.ev_count = pl->d_ev_count;
	gp.ev_cap = pl->ev_cap;
	return gp;
}

static int run_probe(dfb_plan* pl)
{
	dfb_ctx* ctx = pl->ctx;
	const int sm = ctx->prop.multiProcessorCount;
	CK(ctx, cudaMemsetAsync(pl->d_ev_count, 0, sizeof(unsigned long long), ctx->stream));
	for (int c = 0; c < kNumClasses; c++)
	{
		ClassWork& cw = pl->cls[c];
		if (cw.jobs.empty()) continue;
		CK(ctx, cudaMemsetAsync(cw.d_ctrl + 2, 0, sizeof(int), ctx->stream));
		FastParams fp = cw.fp;
		fp.cursor = cw.d_ctrl + 2;
		fp.events = pl->d_events;
		fp.ev_cap = pl->ev_cap;
		CK(ctx, launch_fast(c, MODE_PROBE, fp, sm, ctx->stream, (int)cw.jobs.size()));
		pl->stats.kernel_launches++;
	}
	if (!pl->gen_jobs.empty())
	{
		CK(ctx, cudaMemsetAsync(pl->d_gen_ctrl + 1, 0, sizeof(int), ctx->stream));
		GenParams gp = gen_params(pl, 1);
		dp_generic_kernel<MODE_PROBE><<<pl->gen_grid, 128, 0, ctx->stream>>>(gp);
		CK(ctx, cudaGetLastError());
		pl->stats.kernel_launches++;
	}
	return DFB_OK;
}

extern "C" int dfb_plan_run(dfb_plan* pl)
{
	if (!pl) return DFB_ERR_ARG;
	dfb_ctx* ctx = pl->ctx;
	CK(ctx, cudaSetDevice(ctx->device));
	const int sm = ctx->prop.multiProcessorCount;
	pl->stats.kernel_launches = 0;
	pl->fetched = false;
	pl->run_timed = false;
	if (pl->timing)
	{
		for (int k = 0; k < 3; k++)
			if (!pl->ev[k]) CK(ctx, cudaEventCreate(&pl->ev[k]));
		CK(ctx, cudaEventRecord(pl->ev[0], ctx->stream));
	}
	for (int c = 0; c < kNumClasses; c++)
	{
		ClassWork& cw = pl->cls[c];
		if (cw.jobs.empty()) continue;
		CK(ctx, cudaMemsetAsync(cw.d_ctrl, 0, 4 * sizeof(int), ctx->stream));
		FastParams fp = cw.fp;
		fp.cursor = cw.d_ctrl;
		CK(ctx, launch_fast(c, pl->split ? MODE_SPLIT : MODE_SIMPLE, fp, sm, ctx->stream, (int)cw.jobs.size()));
		pl->stats.kernel_launches++;
	}
	if (!pl->gen_jobs.empty())
	{
		CK(ctx, cudaMemsetAsync(pl->d_gen_ctrl, 0, 4 * sizeof(int), ctx->stream));
		GenParams gp = gen_params(pl, 0);
		if (pl->split)
		{
			dp_generic_kernel<MODE_SPLIT><<<pl->gen_grid, 128, 0, ctx->stream>>>(gp);
			CK(ctx, cudaGetLastError());
			GenReduceParams rp;
			rp.jobs = pl->d_gen_jobs;
			rp.n_tasks = (int)(pl->gen_jobs.size() / 2);
			rp.task_min_score = pl->d_task_min_score;
			rp.min_split = pl->sp.min_split_score;
			rp.rowmax = pl->d_gen_rowmax;
			rp.row_en = pl->d_gen_row_en;
			rp.out_best = pl->d_out;
			rp.probe_flag = pl->d_gen_probe_flag;
			split_reduce_generic_kernel<<<(rp.n_tasks + 127) / 128, 128, 0, ctx->stream>>>(rp);
			CK(ctx, cudaGetLastError());
			pl->stats.kernel_launches += 2;
		}
		else
		{
			dp_generic_kernel<MODE_SIMPLE><<<pl->gen_grid, 128, 0, ctx->stream>>>(gp);
			CK(ctx, cudaGetLastError());
			pl->stats.kernel_launches++;
		}
	}
	if (pl->timing) CK(ctx, cudaEventRecord(pl->ev[1], ctx->stream));
	if (pl->split)
	{
		int rc = run_probe(pl);
		if (rc) return rc;
	}
	if (pl->timing)
	{
		CK(ctx, cudaEventRecord(pl->ev[2], ctx->stream));
		pl->run_timed = true;
	}
	pl->ran = true;
	return DFB_OK;
}

extern "C" int dfb_plan_sync(dfb_plan* pl)
{
	if (!pl) return DFB_ERR_ARG;
	dfb_ctx* ctx = pl->ctx;
	CK(ctx, cudaSetDevice(ctx->device));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	if (pl->run_timed)
	{
		float a = 0, b = 0;
		CK(ctx, cudaEventElapsedTime(&a, pl->ev[0], pl->ev[1]));
		CK(ctx, cudaEventElapsedTime(&b, pl->ev[1], pl->ev[2]));
		pl->stats.ms_sweep = a;
		pl->stats.ms_probe = b;
	}
	return DFB_OK;
}

extern "C" int dfb_plan_set_timing(dfb_plan* pl, int enable)
{
	if (!pl) return DFB_ERR_ARG;
	pl->timing = enable != 0;
	return DFB_OK;
}

extern "C" int dfb_simple_plan_fetch(dfb_plan* pl, int32_t* out_score)
{
	if (!pl) return DFB_ERR_ARG;
	dfb_ctx* ctx = pl->ctx;
	if (pl->split) return set_err(ctx, DFB_ERR_STATE, "dfb_simple_plan_fetch on a split plan");
	if (!pl->ran) return set_err(ctx, DFB_ERR_STATE, "dfb_simple_plan_fetch before dfb_plan_run");
	if (!out_score && pl->n_tasks) return set_err(ctx, DFB_ERR_ARG, "null output");
	CK(ctx, cudaSetDevice(ctx->device));
	if (pl->n_tasks)
		CK(ctx, cudaMemcpyAsync(out_score, pl->d_out, (size_t)pl->n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	pl->stats.d2h_bytes = pl->n_tasks * (int64_t)sizeof(int32_t);
	pl->fetched = true;
	return DFB_OK;
}

extern "C" int dfb_split_plan_fetch(dfb_plan* pl, int32_t* out_best, int64_t* n_rows, int64_t* n_cols)
{
	if (!pl) return DFB_ERR_ARG;
	dfb_ctx* ctx = pl->ctx;
	if (!pl->split) return set_err(ctx, DFB_ERR_STATE, "dfb_split_plan_fetch on a simple plan");
	if (!pl->ran) return set_err(ctx, DFB_ERR_STATE, "dfb_split_plan_fetch before dfb_plan_run");
	CK(ctx, cudaSetDevice(ctx->device));
	unsigned long long n_ev = 0;
	for (int attempt = 0;; attempt++)
	{
		CK(ctx, cudaMemcpyAsync(&n_ev, pl->d_ev_count, sizeof(n_ev), cudaMemcpyDeviceToHost, ctx->stream));
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		if (n_ev <= pl->ev_cap) break;
		if (attempt >= 2) return set_err(ctx, DFB_ERR_STATE, "event buffer overflow persists");
		// the probe sweep found more arg-max columns than the buffer holds: size it exactly and redo the sweep
		dfree(pl->d_events);
		pl->ev_cap = n_ev + 1024;
		cudaError_t e = cudaMalloc(&pl->d_events, (size_t)pl->ev_cap * sizeof(Event));
		if (e != cudaSuccess) return set_err(ctx, DFB_ERR_NOMEM, "event buffer of %llu entries: %s", pl->ev_cap, cudaGetErrorString(e));
		int rc = run_probe(pl);
		if (rc) return rc;
	}
	std::vector<Event> ev((size_t)n_ev);
	if (n_ev) CK(ctx, cudaMemcpyAsync(ev.data(), pl->d_events, (size_t)n_ev * sizeof(Event), cudaMemcpyDeviceToHost, ctx->stream));
	if (out_best && pl->n_tasks)
		CK(ctx, cudaMemcpyAsync(out_best, pl->d_out, (size_t)pl->n_tasks * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	int probe_jobs = 0;
	for (int c = 0; c < kNumClasses; c++)
	{
		if (pl->cls[c].jobs.empty()) continue;
		int h = 0;
		CK(ctx, cudaMemcpyAsync(&h, pl->cls[c].d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		CK(ctx, cudaStreamSynchronize(ctx->stream));
		probe_jobs += h;
	}
	CK(ctx, cudaStreamSynchronize(ctx->stream));
	pl->stats.d2h_bytes = (int64_t)(n_ev * sizeof(Event)) + pl->n_tasks * (int64_t)sizeof(int32_t);
	pl->stats.events = (int64_t)n_ev;
	pl->stats.probe_jobs = probe_jobs;

	// order: task, then matrix (half), then row, then column -- the nested loops of
	// SplitReadAligner.cpp:233-269 walk rows and columns ascending
	std::sort(ev.begin(), ev.end(), [](const Event& a, const Event& b) {
		if (a.task != b.task) return a.task < b.task;
		if (a.half_row != b.half_row) return a.half_row < b.half_row;
		return a.col < b.col;
	});
	pl->rows.clear();
	pl->cols.clear();
	pl->cols.reserve(ev.size());
	size_t i = 0;
	struct Run
	{
		int row, score;
		size_t begin, n;
	};
	std::vector<Run> r1, r2;
	while (i < ev.size())
	{
		const int task = ev[i].task;
		r1.clear();
		r2.clear();
		while (i < ev.size() && ev[i].task == task)
		{
			const int hr = ev[i].half_row;
			Run r;
			r.row = hr & 0x3fffffff;
			r.score = ev[i].score;
			r.begin = i;
			while (i < ev.size() && ev[i].task == task && ev[i].half_row == hr) i++;
			r.n = i - r.begin;
			((hr >> 30) ? r2 : r1).push_back(r);
		}
		const int L = pl->task_L[task];
		for (const Run& a : r1) // ascending read_split
		{
			const int want = L - a.row;
			const Run* b = nullptr;
			for (const Run& c : r2)
				if (c.row == want) { b = &c; break; }
			if (!b) continue;
			dfb_split_row row;
			row.task = task;
			row.read_split = a.row;
			row.score1 = a.score;
			row.score2 = b->score;
			row.col1_begin = (int64_t)pl->cols.size();
			row.n1 = (int32_t)a.n;
			for (size_t k = 0; k < a.n; k++) pl->cols.push_back(ev[a.begin + k].col);
			row.col2_begin = (int64_t)pl->cols.size();
			row.n2 = (int32_t)b->n;
			for (size_t k = 0; k < b->n; k++) pl->cols.push_back(ev[b->begin + k].col);
			pl->rows.push_back(row);
		}
	}
	if (n_rows) *n_rows = (int64_t)pl->rows.size();
	if (n_cols) *n_cols = (int64_t)pl->cols.size();
	pl->fetched = true;
	return DFB_OK;
}

extern "C" int dfb_split_plan_copy(const dfb_plan* pl, dfb_split_row* rows, int32_t* cols)
{
	if (!pl) return DFB_ERR_ARG;
	if (!pl->fetched || !pl->split) return set_err(pl->ctx, DFB_ERR_STATE, "dfb_split_plan_copy before dfb_split_plan_fetch");
	if (rows && !pl->rows.empty()) memcpy(rows, pl->rows.data(), pl->rows.size() * sizeof(dfb_split_row));
	if (cols && !pl->cols.empty()) memcpy(cols, pl->cols.data(), pl->cols.size() * sizeof(int32_t));
	return DFB_OK;
}

extern "C" int dfb_plan_get_stats(const dfb_plan* pl, dfb_plan_stats* stats)
{
	if (!pl || !stats) return DFB_ERR_ARG;
	*stats = pl->stats;
	return DFB_OK;
}

// ---- one-call forms -------------------------------------------------------------------------------

extern "C" int dfb_simple_align_batch(dfb_ctx* ctx, const dfb_simple_params* params, const dfb_seq_table* refs,
                                      const dfb_seq_table* seqs, const int32_t* task_ref, const int32_t* task_seq,
                                      int64_t n_tasks, int32_t* out_score)
{
	dfb_plan* pl = nullptr;
	int rc = dfb_simple_plan_create(ctx, params, refs, seqs, task_ref, task_seq, n_tasks, &pl);
	if (rc) return rc;
	rc = dfb_plan_run(pl);
	if (!rc) rc = dfb_simple_plan_fetch(pl, out_score);
	dfb_plan_destroy(pl);
	return rc;
}

extern "C" int dfb_split_align_batch(dfb_ctx* ctx, const dfb_split_params* params, const dfb_seq_table* refs,
                                     const dfb_seq_table* reads, const int32_t* task_cluster, const int32_t* task_read,
                                     const int32_t* task_min_score, int64_t n_tasks, int32_t* out_best)
{
	if (!ctx) return DFB_ERR_ARG;
	if (ctx->last_split)
	{
		dfb_plan_destroy(ctx->last_split);
		ctx->last_split = nullptr;
	}
	dfb_plan* pl = nullptr;
	int rc = dfb_split_plan_create(ctx, params, refs, reads, task_cluster, task_read, task_min_score, n_tasks, &pl);
	if (rc) return rc;
	rc = dfb_plan_run(pl);
	if (!rc) rc = dfb_split_plan_fetch(pl, out_best, nullptr, nullptr);
	if (rc)
	{
		dfb_plan_destroy(pl);
		return rc;
	}
	// keep only the host-side result; device memory can go
	ctx->last_split = pl;
	return DFB_OK;
}

extern "C" int dfb_split_result_size(const dfb_ctx* ctx, int64_t* n_rows, int64_t* n_cols)
{
	if (!ctx) return DFB_ERR_ARG;
	if (!ctx->last_split) return set_err(ctx, DFB_ERR_STATE, "no split result on this context");
	if (n_rows) *n_rows = (int64_t)ctx->last_split->rows.size();
	if (n_cols) *n_cols = (int64_t)ctx->last_split->cols.size();
	return DFB_OK;
}

extern "C" int dfb_split_result_copy(const dfb_ctx* ctx, dfb_split_row* rows, int32_t* cols)
{
	if (!ctx) return DFB_ERR_ARG;
	if (!ctx->last_split) return set_err(ctx, DFB_ERR_STATE, "no split result on this context");
	return dfb_split_plan_copy(ctx->last_split, rows, cols);
}
