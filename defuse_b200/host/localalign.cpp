// localalign -- drop-in for the reference tool of the same name (tools/localalign.cpp):
//   localalign -m <match> -x <mismatch> -g <gap> [-t <threshold>]  < "id \t reference \t sequence" lines  > "id \t score \t percent"
// Same flags, same input checks and messages, same output bytes and order; SimpleAligner::Align
// (tools/localalign.cpp:79) runs on the GPU in batches through dfb_simple_align_batch.
//
// Host side: stdin arrives in blocks of whole lines (mapped when it is a file, read ahead on a
// helper thread when it is a pipe), a block is parsed in place by line-aligned chunks on all host threads, and the
// 2 001-bp references the pipeline repeats on ~100 consecutive lines (scripts/prep_local_alignment_seqs.pl:120) are
// uploaded once per batch.
//
// Built with -DDFB_COMPACT_INPUT the same source is localalign_dedup: same flags and output, but
// a reference field that is the single character '=' stands for the reference of the line before it.  The pipeline
// repeats each 2 001-bp reference on about a hundred consecutive lines (scripts/prep_local_alignment_seqs.pl:92-156), so
// this form carries every reference once: stdin shrinks from ~2.1 kB to ~120 B per line (INTEGRATION.md shows the
// one-line change on the Perl side).  The stock form stays what it was: a separate binary, not a changed grammar.
#include "host_common.h"
#include "fast_io.h"

#include <string_view>
#include <unordered_map>

using namespace dfbhost;

namespace
{
struct Line
{
	const char* id;
	const char* ref;
	const char* seq;
	uint32_t id_len, ref_len, seq_len;
	int32_t ref_index; // chunk-local while parsing, block-level afterwards
};

struct Chunk
{
	std::vector<Line> lines;
	std::vector<std::string_view> refs; // distinct references of the chunk, in order of first use
	int64_t error_line = -1;            // 0-based line inside the block
	const char* error = nullptr;
};
}  // namespace

int main(int argc, char* argv[])
{
	CommandLine cmd("Local realignment tool", {
	    {'m', "match", "Match Score", true, "int", "", false},
	    {'x', "mismatch", "Mismatch Score", true, "int", "", false},
	    {'g', "gap", "Gap Score", true, "int", "", false},
	    {'t', "threshold", "Percent Perfect Threshold", false, "float", "", false},
	});
	cmd.Parse(argc, argv);
	const int match = cmd.Int('m'), mismatch = cmd.Int('x'), gap = cmd.Int('g');
	const double threshold = cmd.Double('t', 0.0);

	Gpu gpu;
	PhaseTimer timer;
	const dfb_simple_params params{match, mismatch, gap};
	const int T = ToolThreads();

	size_t kBatchTasks = 1u << 20;
	if (const char* e = getenv("DFB_TOOL_BATCH")) kBatchTasks = (size_t)std::max(1, atoi(e)); // tests: force several batches
	const size_t kBatchBytes = 1u << 28;
	size_t block_bytes = (size_t)1 << 29;
	if (const char* e = getenv("DFB_TOOL_BLOCK")) block_bytes = (size_t)std::max(1, atoi(e)); // tests: force several blocks

	LineBlocks input;
	input.Open(0, block_bytes);

	TableBuilder refs, seqs;
	std::vector<int32_t> task_ref, task_seq, score, remap;
	std::vector<std::string> out_parts((size_t)T);
	int64_t lines_before = 0; // lines of the blocks already handled (for the 1-based line numbers of the messages)
#ifdef DFB_COMPACT_INPUT
	std::string carry_store;  // the reference of the last line seen, across chunks and blocks
	std::string_view carry;
	bool have_carry = false;
#endif
	const char* block = nullptr;
	size_t block_len = 0;
	while (input.Next(block, block_len))
	{
		timer.Add("read");
		// ---- parse: one chunk of whole lines per thread, nothing copied ----
		std::vector<LineChunk> chunks = SplitLines(block, block_len, T);
		std::vector<Chunk> parsed(chunks.size());
		ParallelRun((int)chunks.size(), [&](int c) {
			Chunk& out = parsed[(size_t)c];
			const char* a = block + chunks[(size_t)c].begin;
			const char* const end = block + chunks[(size_t)c].end;
			int64_t line_index = chunks[(size_t)c].first_line;
			out.lines.reserve((size_t)(end - a) / 1024 + 16);
			std::unordered_map<std::string_view, int32_t> lookup;
			for (; a < end; line_index++)
			{
				const char* nl = (const char*)memchr(a, '\n', (size_t)(end - a));
				const char* const b = a;
				const char* const e = nl ? nl : end;
				a = nl ? nl + 1 : end;
				if (b == e)
				{
					out.error_line = line_index;
					out.error = "Error: Empty line ";
					return;
				}
				const char* t1 = (const char*)memchr(b, '\t', (size_t)(e - b));
				const char* t2 = t1 ? (const char*)memchr(t1 + 1, '\t', (size_t)(e - t1 - 1)) : nullptr;
				if (!t2)
				{
					out.error_line = line_index;
					out.error = "Error: Format error for line ";
					return;
				}
				const char* t3 = (const char*)memchr(t2 + 1, '\t', (size_t)(e - t2 - 1)); // fields behind the third are ignored
				Line ln;
				ln.id = b;
				ln.id_len = (uint32_t)(t1 - b);
				ln.ref = t1 + 1;
				ln.ref_len = (uint32_t)(t2 - t1 - 1);
				ln.seq = t2 + 1;
				ln.seq_len = (uint32_t)((t3 ? t3 : e) - t2 - 1);
#ifdef DFB_COMPACT_INPUT
				if (ln.ref_len == 1 && ln.ref[0] == '=')
				{
					// the reference of the line before; the first line of a chunk is resolved when the chunks are put together
					if (out.lines.empty()) ln.ref_index = -1;
					else
					{
						ln.ref = out.lines.back().ref;
						ln.ref_len = out.lines.back().ref_len;
						ln.ref_index = out.lines.back().ref_index;
					}
				}
				else
#endif
				if (!out.lines.empty() && out.lines.back().ref_index >= 0 && out.lines.back().ref_len == ln.ref_len &&
				    memcmp(out.lines.back().ref, ln.ref, ln.ref_len) == 0)
				{
					ln.ref_index = out.lines.back().ref_index;
				}
				else
				{
					const std::string_view key(ln.ref, ln.ref_len);
					auto ins = lookup.emplace(key, (int32_t)out.refs.size());
					if (ins.second) out.refs.push_back(key);
					ln.ref_index = ins.first->second;
				}
				out.lines.push_back(ln);
			}
		});
		// the first bad line ends the input: everything in front of it is still aligned and printed (the reference
		// works line by line), then the message, exit 1
		int64_t error_line = -1;
		const char* error = nullptr;
		size_t usable_chunks = parsed.size();
		for (size_t c = 0; c < parsed.size(); c++)
			if (parsed[c].error)
			{
				error_line = parsed[c].error_line;
				error = parsed[c].error;
				usable_chunks = c + 1;
				break;
			}
		// block-level reference indices
		std::vector<std::string_view> block_refs;
		{
			std::unordered_map<std::string_view, int32_t> lookup;
			for (size_t c = 0; c < usable_chunks; c++)
			{
				std::vector<int32_t> to_block(parsed[c].refs.size());
				for (size_t r = 0; r < parsed[c].refs.size(); r++)
				{
					auto ins = lookup.emplace(parsed[c].refs[r], (int32_t)block_refs.size());
					if (ins.second) block_refs.push_back(parsed[c].refs[r]);
					to_block[r] = ins.first->second;
				}
#ifdef DFB_COMPACT_INPUT
				// leading '=' lines of the chunk: the reference of the last line in front of them (of the previous chunk, or
				// of the previous block); with nothing in front of them the line is malformed
				for (Line& ln : parsed[c].lines)
				{
					if (ln.ref_index >= 0) break;
					if (!have_carry)
					{
						error_line = chunks[c].first_line + (int64_t)(&ln - parsed[c].lines.data());
						error = "Error: Format error for line ";
						break;
					}
					ln.ref = carry.data();
					ln.ref_len = (uint32_t)carry.size();
					auto ins = lookup.emplace(carry, (int32_t)block_refs.size());
					if (ins.second) block_refs.push_back(carry);
					ln.ref_index = -2 - ins.first->second; // already a block-level index (encoded: the loop below skips it)
				}
				if (error && error_line >= chunks[c].first_line && !have_carry && !parsed[c].lines.empty() && parsed[c].lines[0].ref_index == -1)
				{
					// keep only the lines in front of the malformed one
					parsed[c].lines.resize((size_t)(error_line - chunks[c].first_line));
					usable_chunks = c + 1;
				}
#endif
				for (Line& ln : parsed[c].lines) ln.ref_index = ln.ref_index <= -2 ? -2 - ln.ref_index : to_block[(size_t)ln.ref_index];
#ifdef DFB_COMPACT_INPUT
				if (!parsed[c].lines.empty())
				{
					carry = std::string_view(parsed[c].lines.back().ref, parsed[c].lines.back().ref_len);
					have_carry = true;
				}
				if (c + 1 == usable_chunks) break;
#endif
			}
		}
		std::vector<size_t> chunk_base(usable_chunks + 1, 0);
		for (size_t c = 0; c < usable_chunks; c++) chunk_base[c + 1] = chunk_base[c] + parsed[c].lines.size();
		const size_t n_lines = chunk_base[usable_chunks];
		auto line_at = [&](size_t k) -> const Line& {
			const size_t c = (size_t)(std::upper_bound(chunk_base.begin(), chunk_base.end(), k) - chunk_base.begin()) - 1;
			return parsed[c].lines[k - chunk_base[c]];
		};
		timer.Add("parse");

		// ---- batches of tasks: tables, GPU, formatted output in input order ----
		remap.assign(block_refs.size(), -1);
		for (size_t first = 0; first < n_lines;)
		{
			refs.Clear();
			seqs.Clear();
			std::vector<int32_t> used;
			size_t last = first;
			size_t bytes = 0;
			for (; last < n_lines && last - first < kBatchTasks && bytes < kBatchBytes; last++)
			{
				const Line& ln = line_at(last);
				if (remap[(size_t)ln.ref_index] < 0)
				{
					remap[(size_t)ln.ref_index] = (int32_t)refs.Add(ln.ref, ln.ref_len);
					used.push_back(ln.ref_index);
					bytes += ln.ref_len;
				}
				seqs.off.push_back(seqs.off.back() + (int64_t)ln.seq_len);
				bytes += ln.seq_len;
			}
			const size_t n = last - first;
			seqs.bytes.resize((size_t)seqs.off.back());
			task_ref.resize(n);
			task_seq.resize(n);
			score.resize(n);
			ParallelRun(T, [&](int tid) {
				for (size_t k = n * (size_t)tid / (size_t)T; k < n * ((size_t)tid + 1) / (size_t)T; k++)
				{
					const Line& ln = line_at(first + k);
					if (ln.seq_len) memcpy(&seqs.bytes[0] + seqs.off[k], ln.seq, ln.seq_len);
					task_ref[k] = remap[(size_t)ln.ref_index];
					task_seq[k] = (int32_t)k;
				}
			});
			for (int32_t r : used) remap[(size_t)r] = -1;
			timer.Add("tables");
			dfb_seq_table rt = refs.View(), st = seqs.View();
			if (dfb_simple_align_batch(gpu.ctx(), &params, &rt, &st, task_ref.data(), task_seq.data(), (int64_t)n, score.data()) != DFB_OK)
				gpu.Die("alignment failed");
			timer.Add("gpu");
			ParallelRun(T, [&](int tid) {
				std::string& os = out_parts[(size_t)tid];
				os.clear();
				char num[64];
				for (size_t k = n * (size_t)tid / (size_t)T; k < n * ((size_t)tid + 1) / (size_t)T; k++)
				{
					const Line& ln = line_at(first + k);
					const int max_score = (int)ln.seq_len * match;                // tools/localalign.cpp:81
					const double percent = (double)score[k] / (double)max_score; // :82  (0/0 prints as -nan, like the reference)
					if (percent < threshold) continue;                           // :84
					os.append(ln.id, ln.id_len);
					os += '\t';
					AppendInt(os, score[k]);
					os += '\t';
					os.append(num, (size_t)snprintf(num, sizeof(num), "%.6g", percent)); // ostream's default float format
					os += '\n';
				}
			});
			for (const std::string& part : out_parts) fwrite(part.data(), 1, part.size(), stdout);
			fflush(stdout);
			timer.Add("format + write");
			first = last;
		}
		if (error)
		{
			std::cerr << error << lines_before + error_line + 1 << std::endl;
			ExitNow(1);
		}
		if (!chunks.empty()) lines_before += chunks.back().first_line + (int64_t)parsed.back().lines.size();
#ifdef DFB_COMPACT_INPUT
		// the block's memory goes away with the next one: keep the last reference (only now: lines of this block may
		// still have pointed at the previous block's copy)
		if (have_carry && carry.data() != carry_store.data())
		{
			std::string keep(carry.data(), carry.size());
			carry_store.swap(keep);
			carry = carry_store;
		}
#endif
	}
	timer.Report();
	FinishProcess(0);
}
