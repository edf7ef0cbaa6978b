// TEST INFRASTRUCTURE ONLY -- a stand-in for the GPU behind the C ABI, for testing the tools' HOST LOGIC without a GPU.
//
// The drop-in tools (defuse_b200/host/*.cpp) are host code around eight C-ABI calls.  Everything around those calls --
// command line, ingest (mapped files, chunked parallel parsing), candidate enumeration, task construction, batching,
// sharding over contexts, record formatting, error exits -- runs on the CPU and is what `-m "not gpu"` tests must
// cover.  This file implements those eight entry points of include/defuse_b200.h by calling the parity oracle
// (oracle/dp_oracle.c, the C restatement of the reference), so that tests/test_tools_host_logic.py can run the real
// tool binaries under LD_PRELOAD on a machine without a GPU and compare their bytes with the compiled reference tools.
//
// It is not part of the product and not a fallback: it is compiled by the test into a temporary directory, never
// installed next to the tools, and nothing in defuse_b200/ knows it exists.  Without LD_PRELOAD the tools load
// libdefuse_b200.so, whose dfb_ctx_create fails when there is no sm_100 GPU (tests/test_tools_cli.py checks that).
#include "defuse_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C"
{
int dpo_simple_align(const uint8_t* ref, int R, const uint8_t* seq, int L, int match, int mismatch, int gap);
int64_t dpo_split_align(const uint8_t* read, int L, const uint8_t* ref1, int R1, const uint8_t* ref2, int R2, int match,
                        int mismatch, int gap, int end_gaps, int min_split_score, int min_score, int32_t* out, int64_t cap,
                        int32_t* rowmax1, int32_t* rowmax2);
void dpo_split_backtrace(const uint8_t* read, int L, const uint8_t* ref1, int R1, const uint8_t* ref2, int R2, int match,
                         int mismatch, int gap, int end_gaps, int ref_split1, int ref_split2, int read_split,
                         int32_t* matches1, int32_t* n1, int32_t* matches2, int32_t* n2);
}

struct dfb_ctx
{
	int device = 0;
	std::string error;
	std::vector<dfb_split_row> rows;
	std::vector<int32_t> cols;
	bool have_result = false;
};

namespace
{
const int kDevices = 4; // what dfb_device_count reports (DFB_DEVICES=all in the tools)
thread_local std::string g_create_error;

template <class F>
void ParallelTasks(int64_t n, F fn)
{
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>(8, n / 16 + 1));
	std::atomic<int64_t> next{0};
	std::vector<std::thread> th;
	auto work = [&] {
		for (int64_t t; (t = next.fetch_add(1)) < n;) fn(t);
	};
	for (int k = 1; k < T; k++) th.emplace_back(work);
	work();
	for (auto& t : th) t.join();
}

struct Seq
{
	const uint8_t* p;
	int n;
};
Seq At(const dfb_seq_table* t, int64_t k) { return Seq{t->bytes + t->off[k], (int)(t->off[k + 1] - t->off[k])}; }

// DEVICE_DOUBLE_SKIP_DP=1: answer every call with "no alignment" at once -- for timing the tools' host phases at
// sizes the oracle would need hours for (scripts/cpu_tool_host_phases.py)
bool SkipDp()
{
	static const bool skip = getenv("DEVICE_DOUBLE_SKIP_DP") && *getenv("DEVICE_DOUBLE_SKIP_DP") == '1';
	return skip;
}

// DEVICE_DOUBLE_DIGEST=<file>: append one line per split/simple call with a hash of everything the tool handed over
// (task order, the window pair and the read bytes of every task, thresholds) -- two builds of a tool submit the same
// work exactly when their digests agree, also at sizes where the DP itself is skipped
void Digest(const char* what, const dfb_seq_table* a, const dfb_seq_table* b, const int32_t* ta, const int32_t* tb,
            const int32_t* extra, int64_t n_tasks, bool pairs)
{
	const char* path = getenv("DEVICE_DOUBLE_DIGEST");
	if (!path || !*path) return;
	uint64_t h = 1469598103934665603ull;
	auto mix = [&](const void* p, size_t n) {
		const uint8_t* q = (const uint8_t*)p;
		for (size_t k = 0; k < n; k++) h = (h ^ q[k]) * 1099511628211ull;
	};
	for (int64_t t = 0; t < n_tasks; t++)
	{
		if (pairs)
		{
			const Seq r1 = At(a, 2 * (int64_t)ta[t]), r2 = At(a, 2 * (int64_t)ta[t] + 1);
			mix(r1.p, (size_t)r1.n);
			mix("|", 1);
			mix(r2.p, (size_t)r2.n);
		}
		else
		{
			const Seq r = At(a, ta[t]);
			mix(r.p, (size_t)r.n);
		}
		mix("|", 1);
		const Seq s = At(b, tb[t]);
		mix(s.p, (size_t)s.n);
		mix("|", 1);
		if (extra) mix(&extra[t], sizeof(int32_t));
	}
	if (FILE* f = fopen(path, "a"))
	{
		fprintf(f, "%s %lld %016llx\n", what, (long long)n_tasks, (unsigned long long)h);
		fclose(f);
	}
}

int Fail(dfb_ctx* ctx, int code, const char* what)
{
	if (ctx) ctx->error = what;
	return code;
}
}  // namespace

extern "C"
{
int dfb_abi_version(void) { return DFB_ABI_VERSION; }
int dfb_device_count(void) { return kDevices; }

int dfb_ctx_create(int device_ordinal, dfb_ctx** out)
{
	if (!out) return DFB_ERR_ARG;
	*out = nullptr;
	if (device_ordinal < 0 || device_ordinal >= kDevices)
	{
		g_create_error = "device double: no such device";
		return DFB_ERR_NODEVICE;
	}
	*out = new dfb_ctx();
	(*out)->device = device_ordinal;
	return DFB_OK;
}

void dfb_ctx_destroy(dfb_ctx* ctx) { delete ctx; }

const char* dfb_last_error(const dfb_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int dfb_simple_align_batch(dfb_ctx* ctx, const dfb_simple_params* params, const dfb_seq_table* refs, const dfb_seq_table* seqs,
                           const int32_t* task_ref, const int32_t* task_seq, int64_t n_tasks, int32_t* out_score)
{
	if (!ctx || !params || !refs || !seqs || (n_tasks > 0 && (!task_ref || !task_seq || !out_score)))
		return Fail(ctx, DFB_ERR_ARG, "device double: null argument");
	for (int64_t t = 0; t < n_tasks; t++)
		if (task_ref[t] < 0 || task_ref[t] >= refs->n || task_seq[t] < 0 || task_seq[t] >= seqs->n)
			return Fail(ctx, DFB_ERR_ARG, "device double: table index out of range");
	Digest("simple", refs, seqs, task_ref, task_seq, nullptr, n_tasks, false);
	if (SkipDp())
	{
		std::fill(out_score, out_score + n_tasks, 0);
		return DFB_OK;
	}
	ParallelTasks(n_tasks, [&](int64_t t) {
		const Seq r = At(refs, task_ref[t]), s = At(seqs, task_seq[t]);
		out_score[t] = dpo_simple_align(r.p, r.n, s.p, s.n, params->match, params->mismatch, params->gap);
	});
	return DFB_OK;
}

// rows in the factorised form of dfb_split_row, rebuilt from the oracle's flat tuple list: tuples come grouped by
// readSplit.first (ascending), inside a group i1-major / i2-minor, both ascending
int dfb_split_align_batch(dfb_ctx* ctx, const dfb_split_params* params, const dfb_seq_table* refs, const dfb_seq_table* reads,
                          const int32_t* task_cluster, const int32_t* task_read, const int32_t* task_min_score, int64_t n_tasks,
                          int32_t* out_best)
{
	if (!ctx || !params || !refs || !reads || (n_tasks > 0 && (!task_cluster || !task_read || !task_min_score)))
		return Fail(ctx, DFB_ERR_ARG, "device double: null argument");
	if (refs->n & 1) return Fail(ctx, DFB_ERR_ARG, "refs must hold two windows per cluster (n is odd)");
	for (int64_t t = 0; t < n_tasks; t++)
		if (task_cluster[t] < 0 || 2 * (int64_t)task_cluster[t] + 1 >= refs->n || task_read[t] < 0 || task_read[t] >= reads->n)
			return Fail(ctx, DFB_ERR_ARG, "device double: table index out of range");
	struct TaskOut
	{
		std::vector<dfb_split_row> rows;
		std::vector<int32_t> cols;
		int best = 0;
	};
	Digest("split", refs, reads, task_cluster, task_read, task_min_score, n_tasks, true);
	if (SkipDp())
	{
		ctx->rows.clear();
		ctx->cols.clear();
		if (out_best) std::fill(out_best, out_best + n_tasks, 0);
		ctx->have_result = true;
		return DFB_OK;
	}
	std::vector<TaskOut> outs((size_t)n_tasks);
	ParallelTasks(n_tasks, [&](int64_t t) {
		const Seq r1 = At(refs, 2 * (int64_t)task_cluster[t]), r2 = At(refs, 2 * (int64_t)task_cluster[t] + 1), rd = At(reads, task_read[t]);
		std::vector<int32_t> rm1((size_t)rd.n + 1), rm2((size_t)rd.n + 1);
		auto run = [&](int32_t* buf, int64_t cap) {
			return dpo_split_align(rd.p, rd.n, r1.p, r1.n, r2.p, r2.n, params->match, params->mismatch, params->gap,
			                       params->end_gaps, params->min_split_score, task_min_score[t], buf, cap, rm1.data(), rm2.data());
		};
		const int64_t n = run(nullptr, 0);
		TaskOut& o = outs[(size_t)t];
		// GetAlignments' maxScore (SplitReadAligner.cpp:194-222): also defined when the winning rows emit nothing
		for (int a = 0; a <= rd.n; a++)
		{
			const int tot = rm1[(size_t)a] + rm2[(size_t)(rd.n - a)];
			if (tot >= task_min_score[t] && tot > o.best) o.best = tot;
		}
		if (n == 0) return;
		std::vector<int32_t> tup((size_t)(7 * n));
		run(tup.data(), n);
		for (int64_t k = 0; k < n;)
		{
			const int32_t a = tup[(size_t)(7 * k + 2)];
			int64_t e = k;
			while (e < n && tup[(size_t)(7 * e + 2)] == a) e++;
			int64_t n2 = 0;
			while (k + n2 < e && tup[(size_t)(7 * (k + n2))] == tup[(size_t)(7 * k)]) n2++;
			const int64_t n1 = (e - k) / n2;
			dfb_split_row row;
			row.task = (int32_t)t;
			row.read_split = a;
			row.score1 = tup[(size_t)(7 * k + 5)];
			row.score2 = tup[(size_t)(7 * k + 6)];
			row.col_begin = (int64_t)o.cols.size();
			row.n1 = (int32_t)n1;
			row.n2 = (int32_t)n2;
			for (int64_t p = 0; p < n1; p++) o.cols.push_back(tup[(size_t)(7 * (k + p * n2))]);
			for (int64_t q = 0; q < n2; q++) o.cols.push_back(r2.n - tup[(size_t)(7 * (k + q) + 1)] - 1);
			o.rows.push_back(row);
			k = e;
		}
	});
	ctx->rows.clear();
	ctx->cols.clear();
	for (int64_t t = 0; t < n_tasks; t++)
	{
		TaskOut& o = outs[(size_t)t];
		if (out_best) out_best[t] = o.best;
		const int64_t shift = (int64_t)ctx->cols.size();
		for (dfb_split_row row : o.rows)
		{
			row.col_begin += shift;
			ctx->rows.push_back(row);
		}
		ctx->cols.insert(ctx->cols.end(), o.cols.begin(), o.cols.end());
	}
	ctx->have_result = true;
	return DFB_OK;
}

int dfb_split_result_size(const dfb_ctx* ctx, int64_t* n_rows, int64_t* n_cols)
{
	if (!ctx || !ctx->have_result) return DFB_ERR_STATE;
	if (n_rows) *n_rows = (int64_t)ctx->rows.size();
	if (n_cols) *n_cols = (int64_t)ctx->cols.size();
	return DFB_OK;
}

int dfb_split_result_copy(const dfb_ctx* ctx, dfb_split_row* rows, int32_t* cols)
{
	if (!ctx || !ctx->have_result) return DFB_ERR_STATE;
	if (rows && !ctx->rows.empty()) memcpy(rows, ctx->rows.data(), ctx->rows.size() * sizeof(dfb_split_row));
	if (cols && !ctx->cols.empty()) memcpy(cols, ctx->cols.data(), ctx->cols.size() * sizeof(int32_t));
	return DFB_OK;
}

int dfb_split_result_view(const dfb_ctx* ctx, const dfb_split_row** rows, int64_t* n_rows, const int32_t** cols, int64_t* n_cols)
{
	if (!ctx || !ctx->have_result) return DFB_ERR_STATE;
	if (rows) *rows = ctx->rows.data();
	if (n_rows) *n_rows = (int64_t)ctx->rows.size();
	if (cols) *cols = ctx->cols.data();
	if (n_cols) *n_cols = (int64_t)ctx->cols.size();
	return DFB_OK;
}

int dfb_split_backtrace_batch(dfb_ctx* ctx, const dfb_split_params* params, const dfb_seq_table* refs, const dfb_seq_table* reads,
                              const int32_t* task_cluster, const int32_t* task_read, const int32_t* task_ref_split1,
                              const int32_t* task_ref_split2, const int32_t* task_read_split, int64_t n_tasks, int64_t* match_off,
                              int32_t* matches, int64_t matches_cap, int64_t* n_pairs)
{
	if (!ctx || !params || !refs || !reads || !match_off) return Fail(ctx, DFB_ERR_ARG, "device double: null argument");
	std::vector<std::vector<int32_t>> m1((size_t)n_tasks), m2((size_t)n_tasks);
	for (int64_t t = 0; t < n_tasks; t++)
	{
		if (task_cluster[t] < 0 || 2 * (int64_t)task_cluster[t] + 1 >= refs->n || task_read[t] < 0 || task_read[t] >= reads->n)
			return Fail(ctx, DFB_ERR_ARG, "device double: table index out of range");
		const Seq r1 = At(refs, 2 * (int64_t)task_cluster[t]), r2 = At(refs, 2 * (int64_t)task_cluster[t] + 1), rd = At(reads, task_read[t]);
		const int a = task_read_split[t], i1 = task_ref_split1[t], i2 = r2.n - task_ref_split2[t] - 1;
		if (a < 0 || a > rd.n || i1 < 0 || i1 > r1.n || i2 < 0 || i2 > r2.n)
			return Fail(ctx, DFB_ERR_ARG, "device double: start cell outside the matrix");
	}
	ParallelTasks(n_tasks, [&](int64_t t) {
		const Seq r1 = At(refs, 2 * (int64_t)task_cluster[t]), r2 = At(refs, 2 * (int64_t)task_cluster[t] + 1), rd = At(reads, task_read[t]);
		m1[(size_t)t].resize((size_t)(2 * rd.n + 2));
		m2[(size_t)t].resize((size_t)(2 * rd.n + 2));
		int32_t n1 = 0, n2 = 0;
		dpo_split_backtrace(rd.p, rd.n, r1.p, r1.n, r2.p, r2.n, params->match, params->mismatch, params->gap, params->end_gaps,
		                    task_ref_split1[t], task_ref_split2[t], task_read_split[t], m1[(size_t)t].data(), &n1,
		                    m2[(size_t)t].data(), &n2);
		m1[(size_t)t].resize((size_t)(2 * n1));
		m2[(size_t)t].resize((size_t)(2 * n2));
	});
	int64_t pairs = 0;
	match_off[0] = 0;
	for (int64_t t = 0; t < n_tasks; t++)
	{
		for (const std::vector<int32_t>* m : {&m1[(size_t)t], &m2[(size_t)t]})
		{
			const int64_t n = (int64_t)m->size() / 2;
			if (matches)
			{
				if (pairs + n > matches_cap) return Fail(ctx, DFB_ERR_ARG, "device double: matches buffer too small");
				if (n) memcpy(matches + 2 * pairs, m->data(), (size_t)(2 * n) * sizeof(int32_t));
			}
			pairs += n;
			match_off[(m == &m1[(size_t)t]) ? 2 * t + 1 : 2 * t + 2] = pairs;
		}
	}
	if (n_pairs) *n_pairs = pairs;
	return DFB_OK;
}
}  // extern "C"
