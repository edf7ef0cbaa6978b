// The evaluation core of evalsplitalign (tools/evalsplitalign.cpp, SplitAlignmentTask::Evaluate tools/SplitAlignment.cpp:
// 484-594) over a buffer of SORTED dosplitalign records: shared by the drop-in evalsplitalign (which maps the sorted
// file) and by dosplitalign_eval, the fused align -> evaluate tool that hands the records over in memory.
#pragma once

#include "fast_io.h"
#include "split_tasks.h"

#include <cmath>

namespace dfbhost
{
// one line of dosplitalign output (SplitAlignment::WriteAlignment, tools/SplitAlignment.cpp:305-317)
struct Record
{
	int fusion_id, fragment, read_end, rev_comp;
	std::pair<int, int> ref_split, read_split;
	int score;
};

// what one region of the records file turns into
struct RegionOutput
{
	std::string seq, brk, pred; // the three output files' bytes
	std::string err;            // stderr, in order
	bool fatal = false;         // err ends with the message the tool dies on
};

// One line of the mapped file: [b, e) without the '\n'; next = start of the following line.
struct LineView
{
	const char *b, *e, *next;
};

inline LineView LineAt(const char* a, const char* file_end)
{
	const char* nl = (const char*)memchr(a, '\n', (size_t)(file_end - a));
	return nl ? LineView{a, nl, nl + 1} : LineView{a, file_end, file_end};
}

// first TAB-separated field of a line as lexical_cast<int> would read it
inline bool FusionIdOf(const LineView& l, int& id)
{
	const char* t = (const char*)memchr(l.b, '\t', (size_t)(l.e - l.b));
	return ParseIntRange(l.b, t ? t : l.e, id);
}

// The first line start at or behind `at` whose fusion id differs from the line in front of it (or `file_end`).  A line
// whose id does not parse is a boundary as well: whoever reads up to it reports it.
inline const char* NextRunBoundary(const char* file_begin, const char* file_end, const char* at)
{
	if (at <= file_begin) return file_begin;
	if (at >= file_end) return file_end;
	// start of the line that contains at-1, i.e. the line in front of the first candidate
	const char* prev = at - 1;
	while (prev > file_begin && prev[-1] != '\n') prev--;
	LineView pl = LineAt(prev, file_end);
	int prev_id = 0;
	bool prev_ok = FusionIdOf(pl, prev_id);
	for (const char* a = pl.next; a < file_end;)
	{
		const LineView l = LineAt(a, file_end);
		int id = 0;
		const bool ok = FusionIdOf(l, id);
		if (!ok || !prev_ok || id != prev_id) return a;
		prev_id = id;
		prev_ok = ok;
		a = l.next;
	}
	return file_end;
}

// SplitAlignment::ReadSortedAlignments (tools/SplitAlignment.cpp:319-370) + SplitAlignmentTask::Evaluate (:484-594)
// + BreakPrediction::Write* (:596-624) over the runs that start in [begin, region_end).  Like the reference's reader,
// a run ends on the first line with another fusion id -- that line is read (and its field count and id are checked)
// before the run in front of it is evaluated, also when it belongs to the next region.
inline void ProcessRegion(const char* begin, const char* region_end, const char* file_end,
                   const std::unordered_map<int, ClusterTask>& tasks, RegionOutput& out)
{
	std::vector<Record> run, support;
	const char* fields[10];
	char num[64];
	auto die = [&](const std::string& message) {
		out.err += message;
		out.fatal = true;
	};
	auto bad_cast = [&](const char* what, const char* b, const char* e) {
		die(std::string("Error: bad lexical cast: ") + what + " '" + std::string(b, e) + "'\n");
	};

	const char* pos = begin;
	LineView pending{nullptr, nullptr, nullptr};
	int pending_id = 0;
	for (;;)
	{
		// ---- one run ----
		run.clear();
		int first_id = -1;
		for (;;)
		{
			LineView l;
			int id;
			if (pending.b)
			{
				if (pending.b >= region_end) break; // the next region's first line: only looked at
				l = pending;
				id = pending_id;
				pending.b = nullptr;
			}
			else
			{
				if (pos >= file_end) break;
				if (run.empty() && pos >= region_end) break;
				l = LineAt(pos, file_end);
				pos = l.next;
				// the reference checks for 7 fields and then reads fields 7 and 8 (:331,359-360); fewer than 9 is an error here
				int nf = 0;
				fields[nf++] = l.b;
				for (const char* t = l.b; nf < 10 && (t = (const char*)memchr(t, '\t', (size_t)(l.e - t))) != nullptr;) fields[nf++] = ++t;
				if (nf < 9)
				{
					die("Error: Format error for candidate reads line:\n" + std::string(l.b, l.e) + "\n");
					return;
				}
				if (!ParseIntRange(fields[0], fields[1] - 1, id))
				{
					bad_cast("fusion id", fields[0], fields[1] - 1);
					return;
				}
				if (!run.empty() && id != first_id)
				{
					pending = l;
					pending_id = id;
					break;
				}
			}
			if (run.empty()) first_id = id;
			// fields of this line (a pending line is cut again: it was only checked)
			int nf = 0;
			fields[nf++] = l.b;
			for (const char* t = l.b; nf < 10 && (t = (const char*)memchr(t, '\t', (size_t)(l.e - t))) != nullptr;) fields[nf++] = ++t;
			const char* const last_end = nf == 10 ? fields[9] - 1 : l.e;
			auto fe = [&](int k) { return k + 1 < nf ? fields[k + 1] - 1 : last_end; };
			Record r;
			r.fusion_id = id;
			static const char* const what[9] = {"fusion id", "fragment index", "read end", "revComp", "ref split", "ref split",
			                                    "read split", "read split", "score"};
			int v[9] = {0};
			for (int k = 1; k < 9; k++)
			{
				if (k == 3)
				{
					// lexical_cast<bool>: exactly "0" or "1"
					if (fe(3) - fields[3] != 1 || (fields[3][0] != '0' && fields[3][0] != '1'))
					{
						bad_cast(what[3], fields[3], fe(3));
						return;
					}
					v[3] = fields[3][0] - '0';
				}
				else if (!ParseIntRange(fields[k], fe(k), v[k]))
				{
					bad_cast(what[k], fields[k], fe(k));
					return;
				}
			}
			r.fragment = v[1];
			r.read_end = v[2];
			r.rev_comp = v[3];
			r.ref_split = std::make_pair(v[4], v[5]);
			r.read_split = std::make_pair(v[6], v[7]);
			r.score = v[8];
			run.push_back(r);
		}
		if (run.empty()) return;

		const int fusion_id = first_id;
		// a fusion id without regions: the reference default-constructs a task here (operator[], tools/evalsplitalign.cpp:102)
		// whose windows are empty -- any split then trips the DebugChecks of Evaluate (:545-546), and a run without a
		// split still gets its empty prediction
		static const ClusterTask no_regions;
		auto it = tasks.find(fusion_id);
		const ClusterTask& task = it == tasks.end() ? no_regions : it->second;
		const int task_id = it == tasks.end() ? 0 : fusion_id; // what the prediction is labelled with: the task's own id (:596-612)

		// ---- SplitAlignmentTask::Evaluate (tools/SplitAlignment.cpp:484-594) ----
		// (a fresh container per fusion: its bucket count, hence its iteration order, must be the reference's)
		std::unordered_map<std::pair<int, int>, int, PairHash> split_score;
		for (const Record& r : run) split_score.insert(std::make_pair(r.ref_split, 0)).first->second += r.score;
		int max_score = -1;
		std::pair<int, int> best(0, 0);
		for (const auto& kv : split_score)
			if (kv.second > max_score)
			{
				best = kv.first;
				max_score = kv.second;
			}
		int count = 0;
		double pos_avg = -1.0, min_avg = -1.0;
		int break_pos[2] = {0, 0};
		support.clear();
		AppendInt(out.seq, task_id);
		out.seq += '\t';
		if (max_score == -1)
		{
			// (every summed score below 0: the reference reports it and writes the empty prediction, break
			// positions uninitialised there, 0 here)
			out.err += "Error: Unable to find max score split\n";
			out.seq += 'N';
		}
		else
		{
			for (const Record& r : run)
				if (r.ref_split == best) support.push_back(r);
			if (!(best.first >= 0 && (size_t)best.first <= task.window[0].length()) ||
			    !(best.second + 1 >= 0 && (size_t)(best.second + 1) < task.window[1].length()))
			{
				out.seq.resize(out.seq.size() - 1 - std::to_string(task_id).size());
				die("Error: split outside the breakpoint windows of fusion " + std::to_string(fusion_id) + "\n");
				return;
			}
			out.seq += task.remainder[0];
			out.seq.append(task.window[0], 0, (size_t)best.first);
			out.seq += '|';
			out.seq.append(task.window[1], (size_t)(best.second + 1), std::string::npos);
			out.seq += task.remainder[1];
			break_pos[0] = task.seq_strand[0] == kPlus ? task.seq_start[0] + best.first - 1
			                                           : task.seq_start[0] + task.seq_length[0] - best.first;
			break_pos[1] = task.seq_strand[1] == kPlus ? task.seq_start[1] + best.second + 1
			                                           : task.seq_start[1] + task.seq_length[1] - best.second - 2;
			double pos_sum = 0.0, min_sum = 0.0;
			for (const Record& r : support)
			{
				const int left = r.read_split.first, right = r.read_split.second;
				const double pos_range = (double)(left + right - 2 * kMinAnchor);
				const double pos_value = std::max(0, left - kMinAnchor);
				const double min_range = floor(0.5 * (double)(left + right - 2 * kMinAnchor));
				const double min_value = std::max(0, std::min(left - kMinAnchor, right - kMinAnchor));
				pos_sum += pos_value / pos_range;
				min_sum += min_value / min_range;
			}
			count = (int)support.size();
			pos_avg = pos_sum / (double)support.size();
			min_avg = min_sum / support.size();
		}

		// BreakPrediction::WriteSequence / WriteBreak / WriteAlignments (:596-624); doubles as operator<< prints them
		out.seq += "\t0\t";
		AppendInt(out.seq, count);
		out.seq += '\t';
		out.seq.append(num, (size_t)snprintf(num, sizeof(num), "%g", pos_avg));
		out.seq += '\t';
		out.seq.append(num, (size_t)snprintf(num, sizeof(num), "%g", min_avg));
		out.seq += '\n';
		for (int end = 0; end <= 1; end++)
		{
			AppendInt(out.brk, task_id);
			out.brk += '\t';
			AppendInt(out.brk, end);
			out.brk += '\t';
			out.brk += task.align_ref_name[end];
			out.brk += task.align_strand[end] == kPlus ? "\t+\t" : "\t-\t";
			AppendInt(out.brk, break_pos[end]);
			out.brk += '\n';
		}
		for (const Record& r : support)
		{
			const int f[9] = {r.fusion_id, r.fragment, r.read_end, r.rev_comp, r.ref_split.first, r.ref_split.second,
			                  r.read_split.first, r.read_split.second, r.score};
			for (int k = 0; k < 9; k++)
			{
				AppendInt(out.pred, f[k]);
				out.pred += '\t';
			}
			out.pred += '\n';
		}
	}
}

// The records in [file_begin, file_end) -- sorted the way the pipeline sorts them (`sort -n -k 1`, scripts/defuse_run.pl:
// 528,533) -- evaluated fusion by fusion; the three outputs are written in record order.  Blocks (bounded memory for the
// formatted output) are cut into one region per thread; blocks and regions begin where the fusion id changes.  Returns
// false after a fatal record (the message has been printed, the outputs hold everything in front of it).
inline bool EvaluateSortedRecords(const char* file_begin, const char* file_end, const std::unordered_map<int, ClusterTask>& tasks,
                                  std::ostream& seq_file, std::ostream& break_file, std::ostream& pred_file)
{
	const int T = ToolThreads();
	size_t block_bytes = (size_t)256 << 20, region_min = (size_t)1 << 16;
	if (const char* e = getenv("DFB_TOOL_CHUNK_MIN")) // tests: regions of a few lines, several blocks
	{
		region_min = (size_t)std::max(1, atoi(e));
		block_bytes = region_min * (size_t)T * 3;
	}
	for (const char* block = file_begin; block < file_end;)
	{
		const char* const block_end =
		    (size_t)(file_end - block) <= block_bytes ? file_end : NextRunBoundary(file_begin, file_end, block + block_bytes);
		const size_t n = (size_t)(block_end - block);
		const int R = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, n / region_min + 1));
		std::vector<const char*> cut((size_t)R + 1, block_end);
		cut[0] = block;
		for (int k = 1; k < R; k++)
			cut[(size_t)k] = std::min(block_end, std::max(cut[(size_t)k - 1], NextRunBoundary(file_begin, file_end, block + n / (size_t)R * (size_t)k)));
		std::vector<RegionOutput> outs((size_t)R);
		ParallelRun(R, [&](int k) {
			if (cut[(size_t)k] < cut[(size_t)k + 1]) ProcessRegion(cut[(size_t)k], cut[(size_t)k + 1], file_end, tasks, outs[(size_t)k]);
		});
		for (const RegionOutput& o : outs)
		{
			seq_file.write(o.seq.data(), (std::streamsize)o.seq.size());
			break_file.write(o.brk.data(), (std::streamsize)o.brk.size());
			pred_file.write(o.pred.data(), (std::streamsize)o.pred.size());
			std::cerr << o.err;
			if (o.fatal)
			{
				seq_file.flush();
				break_file.flush();
				pred_file.flush();
				return false;
			}
		}
		block = block_end;
	}
	return true;
}

}  // namespace dfbhost
