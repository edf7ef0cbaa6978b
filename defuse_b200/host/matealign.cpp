// matealign -- drop-in for the reference tool of the same name (tools/matealign.cpp):
//   matealign -m -x -g [-t] -s <searchlength> -r <reference.fa> -1 <reads1.fastq> -2 <reads2.fastq>  < SAM  > "fragment \t score \t percent"
// For every read whose OTHER end has alignments in the SAM, the read is aligned against the
// searchlength+1 window downstream of each of those alignments (N-padded at the sequence ends),
// in fastq order x SAM order (tools/matealign.cpp:179-223).  SimpleAligner::Align (:209) runs on
// the GPU in batches; everything else is the same host logic, written for batching: the SAM is parsed by
// line-aligned chunks on all host threads, both fastq files are indexed in place meanwhile, windows are cut and
// reverse-complemented in parallel straight into the upload table, records are formatted in parallel.
#include "host_common.h"
#include "fast_io.h"

#include <sstream>

#include <fstream>
#include <string_view>
#include <unordered_map>

using namespace dfbhost;

namespace
{
struct MatePosition
{
	int ref_index;
	int strand;   // 0 plus, 1 minus
	int position; // plus: alignment start; minus: alignment end (1-based)
};

// whole FASTA in memory, id = the full header line after '>' (tools/Sequences.cpp:18-58)
struct FastaSequences
{
	std::unordered_map<std::string, std::string> seqs;
	void Read(const std::string& filename)
	{
		MappedInput in;
		if (!in.OpenFile(filename))
		{
			std::cerr << "Error: unable to open file " << filename << std::endl;
			ExitNow(1);
		}
		const char* a = in.data();
		const char* const end = a + in.size();
		std::string id, sequence;
		while (a < end)
		{
			const char* nl = (const char*)memchr(a, '\n', (size_t)(end - a));
			const char* const e = nl ? nl : end;
			if (e > a)
			{
				if (*a == '>')
				{
					if (!id.empty()) seqs[id].swap(sequence);
					id.assign(a + 1, e);
					sequence.clear();
				}
				else
				{
					sequence.append(a, e);
				}
			}
			a = nl ? nl + 1 : end;
		}
		if (!id.empty()) seqs[id].swap(sequence);
	}
	const std::string* Find(const std::string& id) const
	{
		auto it = seqs.find(id);
		return it == seqs.end() ? nullptr : &it->second;
	}
};

// [start, end] 1-based inclusive of `full`, 'N' where the window leaves the sequence (tools/Sequences.cpp:60-79):
// string(prepend,'N') + full.substr(seq_start-1, seq_length) + string(append,'N'), with substr's own clipping.
struct WindowPlan
{
	long long prepend, from, take, append; // take < 0: substr(pos, huge) = the rest of the sequence
	bool bad;                              // the reference would throw (and abort) building this string
	long long Length(long long full_len) const { return prepend + (take < 0 ? full_len - from : take) + append; }
};

inline WindowPlan PlanWindow(long long full_len, int start, int end)
{
	WindowPlan w;
	const long long seq_start = std::max<long long>(1, start);
	w.prepend = seq_start - start;
	const long long seq_end = std::min<long long>(full_len, end);
	w.append = (long long)end - seq_end;
	w.from = seq_start - 1;
	w.bad = w.from > full_len || w.append < 0;
	w.take = seq_end - seq_start + 1;
	if (!w.bad && w.take > full_len - w.from) w.take = full_len - w.from;
	return w;
}

struct SamEntry
{
	int read_id;
	int strand;
	int position;
	std::string_view ref_name;
};

struct Task
{
	uint32_t record;  // index into the fastq order of the file being walked
	uint32_t mate;    // index into the sorted SAM entries
};
}  // namespace

int main(int argc, char* argv[])
{
	CommandLine cmd("Mate Realignment Tool", {
	    {'m', "match", "Match Score", true, "int", "", false},
	    {'x', "mismatch", "Mismatch Score", true, "int", "", false},
	    {'g', "gap", "Gap Score", true, "int", "", false},
	    {'t', "threshold", "Percent Perfect Threshold", false, "float", "", false},
	    {'s', "searchlength", "Search Length", true, "integer", "", false},
	    {'r', "reference", "Reference Sequences Fasta", true, "string", "", false},
	    {'1', "seq1", "End 1 Sequences", true, "string", "", false},
	    {'2', "seq2", "End 2 Sequences", true, "string", "", false},
	});
	cmd.Parse(argc, argv);
	const int match = cmd.Int('m'), mismatch = cmd.Int('x'), gap = cmd.Int('g');
	const double threshold = cmd.Double('t', 0.0);
	const int search_length = cmd.Int('s');
	const std::string reference_fasta = cmd.Str('r');
	const std::string reads_filename[2] = {cmd.Str('1'), cmd.Str('2')};

	Gpu gpu;
	PhaseTimer timer;
	const dfb_simple_params params{match, mismatch, gap};
	const int T = ToolThreads();

	// ---- SAM on stdin -> alignments per read id, in input order (tools/matealign.cpp:80-158) ----
	MappedInput sam;
	sam.OpenStdin();
	std::vector<SamEntry> entries;
	{
		const char* const base = sam.data();
		std::vector<LineChunk> chunks = SplitLines(base, sam.size(), T);
		struct Part
		{
			std::vector<SamEntry> entries;
			int64_t error_line = -1;
			std::string error;
		};
		std::vector<Part> parts(chunks.size());
		ParallelRun((int)chunks.size(), [&](int c) {
			Part& part = parts[(size_t)c];
			const char* a = base + chunks[(size_t)c].begin;
			const char* const end = base + chunks[(size_t)c].end;
			int64_t line_number = chunks[(size_t)c].first_line;
			auto fail = [&](const std::string& msg) {
				part.error_line = line_number;
				part.error = msg;
			};
			while (a < end)
			{
				const char* nl = (const char*)memchr(a, '\n', (size_t)(end - a));
				const char* const b = a;
				const char* const e = nl ? nl : end;
				a = nl ? nl + 1 : end;
				line_number++;
				if (b == e) return fail("Error: Empty alignment line " + std::to_string(line_number));
				if (*b == '@') continue;
				const char* f[11];
				int nf = 0;
				f[0] = b;
				for (const char* q = b; nf < 10;)
				{
					const char* t = (const char*)memchr(q, '\t', (size_t)(e - q));
					if (!t)
					{
						if (nf == 9) f[++nf] = e + 1;
						break;
					}
					f[++nf] = t + 1;
					q = t + 1;
				}
				if (nf < 10) return fail("Error: Format error for alignment line " + std::to_string(line_number));
				int flag = 0, pos = 0;
				if (!ParseIntRange(f[1], f[2] - 1, flag)) return fail("Error: bad lexical cast: flag '" + std::string(f[1], f[2] - 1) + "'");
				if (!ParseIntRange(f[3], f[4] - 1, pos)) return fail("Error: bad lexical cast: pos '" + std::string(f[3], f[4] - 1) + "'");
				if (f[3] - 1 - f[2] == 1 && *f[2] == '*') continue;
				const char* const qb = f[0];
				const char* const qe = f[1] - 1;
				const char* slash = (const char*)memchr(qb, '/', (size_t)(qe - qb));
				if (!slash || memchr(slash + 1, '/', (size_t)(qe - slash - 1)) || qe - slash != 2 || (slash[1] != '1' && slash[1] != '2'))
					return fail("Error: Unable to interpret qname for alignment line " + std::to_string(line_number));
				int fragment_index = 0;
				if (!ParseIntRange(qb, slash, fragment_index))
					return fail("Error: bad lexical cast: fragment index '" + std::string(qb, slash) + "'");
				SamEntry en;
				en.strand = (flag & 0x0010) == 0 ? 0 : 1;
				const int start = pos, stop = start + (int)(f[10] - 1 - f[9]) - 1;
				en.position = en.strand == 0 ? start : stop;
				en.read_id = PackId(fragment_index, slash[1] == '1' ? 0 : 1);
				en.ref_name = std::string_view(f[2], (size_t)(f[3] - 1 - f[2]));
				part.entries.push_back(en);
			}
		});
		for (const Part& part : parts)
			if (part.error_line >= 0)
			{
				std::cerr << part.error << std::endl;
				ExitNow(1);
			}
		size_t total = 0;
		for (const Part& part : parts) total += part.entries.size();
		entries.reserve(total);
		for (const Part& part : parts) entries.insert(entries.end(), part.entries.begin(), part.entries.end());
	}
	// alignments of one read stay in input order: stable sort by read id, then ranges by binary search
	std::stable_sort(entries.begin(), entries.end(), [](const SamEntry& a, const SamEntry& b) { return (unsigned)a.read_id < (unsigned)b.read_id; });
	std::cerr << "Read alignments" << std::endl;
	timer.Lap("sam");

	FastaSequences reference;
	reference.Read(reference_fasta);
	std::cerr << "Read reference fasta" << std::endl;
	timer.Lap("fasta");
	// reference name -> sequence, resolved once per distinct name
	std::unordered_map<std::string_view, const std::string*> ref_of;
	for (const SamEntry& en : entries)
		if (ref_of.find(en.ref_name) == ref_of.end()) ref_of.emplace(en.ref_name, reference.Find(std::string(en.ref_name)));

	FastqIndex fastq[2];
	const bool ok0 = fastq[0].Open(reads_filename[0]);
	const bool ok1 = fastq[1].Open(reads_filename[1]);
	std::cerr << fastq[0].Message() << fastq[1].Message();
	if (!ok0 || !ok1)
	{
		std::cout << "Error: unable to read sequences" << std::endl;
		ExitNow(1);
	}

	size_t kBatchTasks = 1u << 19;
	if (const char* e = getenv("DFB_TOOL_BATCH")) kBatchTasks = (size_t)std::max(1, atoi(e)); // tests: force several batches
	const size_t kBatchBytes = (size_t)1 << 29;
	// the batch's tables (CSR; the byte arrays are plain arrays so that the worker threads touch them first)
	std::unique_ptr<char[]> window_bytes, read_bytes;
	size_t window_cap = 0, read_cap = 0;
	std::vector<int64_t> window_off, read_off;
	std::vector<const std::string*> full;
	std::vector<WindowPlan> plan;
	std::vector<int32_t> task_ref, task_seq, score;
	std::vector<Task> tasks;
	std::vector<std::string> out_parts((size_t)T);
	static unsigned char complement[256];
	for (int k = 0; k < 256; k++) complement[k] = (unsigned char)k;
	for (int k = 0; k < 8; k++) complement[(unsigned char)"ACGTacgt"[k]] = (unsigned char)"TGCAtgca"[k]; // tools/Common.cpp:32-54

	// ---- tasks: (window, read) in fastq order x SAM order; flushed in batches ----
	for (int file = 0; file <= 1; file++)
	{
		fastq[file].Scan(T, true);
		timer.Add("fastq index");
		const FastqIndex& fq = fastq[file];
		// one batch: tasks [tasks of records r0..r1)
		auto flush = [&]() {
			size_t n = tasks.size();
			if (n == 0) return;
			std::string fault; // what a task-by-task reader dies on; the tasks in front of that one are still aligned and printed
			// window lengths and read slots (a read is uploaded once, its tasks are consecutive): planned by ranges of
			// tasks, offsets from the ranges' sums, bytes written by the same ranges into arrays that keep their size
			// from batch to batch (nothing is zeroed first: every byte in use is written here)
			task_ref.resize(n);
			task_seq.resize(n);
			score.resize(n);
			full.resize(n);
			plan.resize(n);
			if (window_off.size() < n + 1) window_off.resize(n + 1);
			if (read_off.size() < n + 1) read_off.resize(n + 1);
			struct Range
			{
				int64_t window_bytes = 0, read_bytes = 0, reads = 0, missing = -1, bad = -1;
			};
			std::vector<Range> range;
			for (;;)
			{
				range.assign((size_t)T + 1, Range());
				ParallelRun(T, [&](int tid) {
					Range& rg = range[(size_t)tid + 1];
					for (size_t k = n * (size_t)tid / (size_t)T; k < n * ((size_t)tid + 1) / (size_t)T; k++)
					{
						const SamEntry& en = entries[tasks[k].mate];
						const std::string* seq = ref_of.find(en.ref_name)->second;
						full[k] = seq;
						if (!seq)
						{
							if (rg.missing < 0) rg.missing = (int64_t)k;
							continue;
						}
						plan[k] = en.strand == 0 ? PlanWindow((long long)seq->size(), en.position, en.position + search_length)
						                         : PlanWindow((long long)seq->size(), en.position - search_length, en.position);
						if (plan[k].bad && rg.bad < 0) rg.bad = (int64_t)k;
						rg.window_bytes += plan[k].bad ? 0 : plan[k].Length((long long)seq->size());
						if (k == 0 || tasks[k].record != tasks[k - 1].record)
						{
							int id;
							const char* rs;
							uint32_t len;
							fq.Record(tasks[k].record, id, rs, len);
							rg.reads++;
							rg.read_bytes += (int64_t)len;
						}
					}
				});
				int64_t first = -1;
				for (int r = 1; r <= T; r++)
					for (int64_t k : {range[(size_t)r].missing, range[(size_t)r].bad})
						if (k >= 0 && (first < 0 || k < first)) first = k;
				if (first < 0) break;
				const SamEntry& en = entries[tasks[(size_t)first].mate];
				std::ostringstream msg;
				if (!full[(size_t)first])
					msg << "Error: Unable to find sequence " << en.ref_name << std::endl;
				else
					// std::string(n,'N') / substr would throw here in the reference (window starts beyond the sequence, or ends
					// before position 1): it aborts; we report and fail the same way (non-zero exit)
					msg << "Error: window around " << en.position << " outside sequence " << en.ref_name << std::endl;
				fault = msg.str();
				n = (size_t)first;
				tasks.resize(n);
				if (n == 0) break;
			}
			if (n > 0)
			{
				for (int r = 1; r <= T; r++)
				{
					range[(size_t)r].window_bytes += range[(size_t)r - 1].window_bytes;
					range[(size_t)r].read_bytes += range[(size_t)r - 1].read_bytes;
					range[(size_t)r].reads += range[(size_t)r - 1].reads;
				}
				const size_t n_reads = (size_t)range[(size_t)T].reads;
				auto grow = [](std::unique_ptr<char[]>& buf, size_t& cap, size_t need) {
					if (cap >= need + 1) return;
					cap = need + need / 8 + 64;
					buf.reset(new char[cap]);
				};
				grow(window_bytes, window_cap, (size_t)range[(size_t)T].window_bytes);
				grow(read_bytes, read_cap, (size_t)range[(size_t)T].read_bytes);
				window_off[0] = read_off[0] = 0;
				ParallelRun(T, [&](int tid) {
					int64_t w_at = range[(size_t)tid].window_bytes, r_at = range[(size_t)tid].read_bytes, r_idx = range[(size_t)tid].reads;
					for (size_t k = n * (size_t)tid / (size_t)T; k < n * ((size_t)tid + 1) / (size_t)T; k++)
					{
						if (k == 0 || tasks[k].record != tasks[k - 1].record)
						{
							int id;
							const char* rs;
							uint32_t rlen;
							fq.Record(tasks[k].record, id, rs, rlen);
							if (rlen) memcpy(read_bytes.get() + r_at, rs, rlen);
							r_at += (int64_t)rlen;
							read_off[(size_t)++r_idx] = r_at;
						}
						task_ref[k] = (int32_t)k;
						task_seq[k] = (int32_t)r_idx - 1; // (a range that starts inside a read's run uses the read its predecessor stored)
						const WindowPlan& w = plan[k];
						const std::string& seq = *full[k];
						char* dst = window_bytes.get() + w_at;
						const long long take = w.take < 0 ? (long long)seq.size() - w.from : w.take;
						const long long len = w.prepend + take + w.append;
						w_at += len;
						window_off[k + 1] = w_at;
						if (entries[tasks[k].mate].strand == 0)
						{
							// plus-strand mate: the window is reverse-complemented (tools/matealign.cpp:197-201)
							char* q = dst + len;
							for (long long j = 0; j < w.prepend; j++) *--q = 'N';
							const char* src = seq.data() + w.from;
							for (long long j = 0; j < take; j++) *--q = (char)complement[(unsigned char)src[j]];
							for (long long j = 0; j < w.append; j++) *--q = 'N';
						}
						else
						{
							memset(dst, 'N', (size_t)w.prepend);
							if (take) memcpy(dst + w.prepend, seq.data() + w.from, (size_t)take);
							memset(dst + w.prepend + take, 'N', (size_t)w.append);
						}
					}
				});
				timer.Add("tables");
				const dfb_seq_table wt{(const uint8_t*)window_bytes.get(), window_off.data(), (int64_t)n};
				const dfb_seq_table rt{(const uint8_t*)read_bytes.get(), read_off.data(), (int64_t)n_reads};
				if (dfb_simple_align_batch(gpu.ctx(), &params, &wt, &rt, task_ref.data(), task_seq.data(), (int64_t)n, score.data()) != DFB_OK)
					gpu.Die("alignment failed");
				timer.Add("gpu");
				ParallelRun(T, [&](int tid) {
					std::string& os = out_parts[(size_t)tid];
					os.clear();
					char num[64];
					for (size_t k = n * (size_t)tid / (size_t)T; k < n * ((size_t)tid + 1) / (size_t)T; k++)
					{
						int id;
						const char* s;
						uint32_t len;
						fq.Record(tasks[k].record, id, s, len);
						const int max_score = (int)len * match;                       // tools/matealign.cpp:211
						const double percent = (double)score[k] / (double)max_score; // :212
						if (percent < threshold) continue;
						AppendInt(os, IdIndex(id)); // readID.fragmentIndex is a 31-bit field
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
			}
			if (!fault.empty())
			{
				std::cerr << fault;
				ExitNow(1);
			}
			tasks.clear();
		};
		size_t batch_bytes = 0;
		for (size_t r = 0; r < fq.Count(); r++)
		{
			int id;
			const char* s;
			uint32_t len;
			fq.Record(r, id, s, len);
			const unsigned other = (unsigned)PackId(IdIndex(id), 1 - IdEnd(id));
			auto lo = std::lower_bound(entries.begin(), entries.end(), other,
			                           [](const SamEntry& a, unsigned key) { return (unsigned)a.read_id < key; });
			for (; lo != entries.end() && (unsigned)lo->read_id == other; ++lo)
			{
				tasks.push_back(Task{(uint32_t)r, (uint32_t)(lo - entries.begin())});
				batch_bytes += (size_t)search_length + 1 + len;
			}
			if (tasks.size() >= kBatchTasks || batch_bytes >= kBatchBytes)
			{
				flush();
				batch_bytes = 0;
			}
		}
		flush();
		// a malformed record ends this file's stream with the reference's message; a fragment name that is not an
		// integer is where the reference dies
		std::cerr << fastq[file].Message();
		if (fastq[file].Fatal()) ExitNow(1);
	}
	timer.Report();
	FinishProcess(0);
}
