// splitseq -- drop-in for the reference tool of the same name (tools/splitseq.cpp): prints, per fusion, the two
// breakpoint windows and under them every supporting split read laid out base by base along its alignment.
//   splitseq -f ref.fa -e exons.regions -u <frag mean> -s <frag sd> -n <minread> -x <maxread>
//            -r clusters.regions -p <reads prefix> -a splitreads.predalign [-i <fusion id>]
// Reads come from the indexed fastq pair <prefix>.{1,2}.fastq through <prefix>.fqi (tools/ReadIndex.cpp:19-129).
// SplitAlignmentTask::ReAlign (tools/SplitAlignment.cpp:443-464) = Align + GetAlignments(backtrace=true) per
// record; here every record of the file is one task of two batched GPU calls: dfb_split_align_batch (which
// alignments exist, in the reference's emission order) and dfb_split_backtrace_batch (the match lists of the
// alignment whose refSplit the record names).  The text layout (SplitAlignment.cpp:402-438,626-635) is host work.
#include "split_tasks.h"

using namespace dfbhost;

namespace
{
struct Record
{
	int fusion_id, fragment, read_end, rev_comp;
	std::pair<int, int> ref_split;
};

// ReadIndex (tools/ReadIndex.cpp): .fqi = little-endian int64 file offsets, entry (2*fragment + end)
class ReadIndexFile
{
public:
	void Open(const std::string& prefix)
	{
		mIndexName = prefix + ".fqi";
		mIndex.open(mIndexName.c_str(), std::fstream::in | std::fstream::binary);
		if (!mIndex.good())
		{
			std::cerr << "Error: Unable to open file " << mIndexName << std::endl;
			ExitNow(1);
		}
		for (int end = 0; end <= 1; end++)
		{
			mFastqName[end] = prefix + (end == 0 ? ".1.fastq" : ".2.fastq");
			mFastq[end].open(mFastqName[end].c_str(), std::fstream::in | std::fstream::binary);
			if (!mFastq[end].good())
			{
				std::cerr << "Error: Unable to open file " << mFastqName[end] << std::endl;
				ExitNow(1);
			}
		}
	}
	void Find(int fragment, int end, std::string& sequence)
	{
		const long pos = (long)fragment * 2 * (long)sizeof(long) + (long)end * (long)sizeof(long);
		mIndex.seekg(pos);
		IndexFail(fragment);
		long fastq_pos = 0;
		mIndex.read((char*)&fastq_pos, sizeof(long));
		IndexFail(fragment);
		mFastq[end].seekg(fastq_pos);
		FastqFail(fragment, end);
		std::string line[4];
		for (int k = 0; k < 4; k++)
		{
			std::getline(mFastq[end], line[k]);
			FastqFail(fragment, end);
		}
		const std::string::size_type slash = line[0].find_first_of('/');
		if (line[0].empty() || line[0][0] != '@' || slash == std::string::npos || slash + 1 >= line[0].length())
		{
			std::cerr << "Error: Unable to interpret read name " << line[0] << " when searching for fragment " << fragment
			          << " end " << end << std::endl;
			ExitNow(1);
		}
		const char end_name = line[0][slash + 1];
		if (end_name != '1' && end_name != '2')
		{
			std::cerr << "Error: Unable to interpret read end " << line[0] << " when searching for fragment " << fragment
			          << " end " << end << std::endl;
			ExitNow(1);
		}
		const int found_fragment = IntOrDie(line[0].substr(1, slash - 1), "fragment index");
		if (found_fragment != fragment)
		{
			std::cerr << "Error: Fragment index mismatch when interpreting " << line[0] << " and searching for fragment "
			          << fragment << " end " << end << std::endl;
			ExitNow(1);
		}
		if ((end_name == '1' ? 0 : 1) != end)
		{
			std::cerr << "Error: Read end mismatch when interpreting " << line[0] << " and searching for fragment " << fragment
			          << " end " << end << std::endl;
			ExitNow(1);
		}
		sequence = line[1];
	}

private:
	void IndexFail(int fragment)
	{
		if (mIndex.fail())
		{
			std::cerr << "Error: Failure reading index file " << mIndexName << " when searching for fragment " << fragment << std::endl;
			ExitNow(1);
		}
	}
	void FastqFail(int fragment, int end)
	{
		if (mFastq[end].fail())
		{
			std::cerr << "Error: Failure reading fastq file " << mFastqName[end] << " when searching for fragment " << fragment
			          << " end " << end << std::endl;
			ExitNow(1);
		}
	}
	std::string mIndexName, mFastqName[2];
	std::ifstream mIndex, mFastq[2];
};
}  // namespace

int main(int argc, char* argv[])
{
	CommandLine cmd("Fusion sequence prediction by split reads", {
	    {'f', "fasta", "Reference Fasta", true, "string", "", false},
	    {'e', "exons", "Exon Regions Filename", true, "string", "", false},
	    {'u', "ufrag", "Fragment Length Mean", true, "float", "", false},
	    {'s', "sfrag", "Fragment Length Standard Deviation", true, "float", "", false},
	    {'n', "minread", "Minimum Read Length", true, "integer", "", false},
	    {'x', "maxread", "Maximum Read Length", true, "integer", "", false},
	    {'r', "regions", "Fusion Regions Filename", true, "string", "", false},
	    {'p', "prefix", "Reads Filename Prefix", true, "string", "", false},
	    {'a', "align", "Split Alignments Filename", true, "string", "", false},
	    {'i', "id", "Query Fusion ID", false, "integer", "-1", false},
	});
	cmd.Parse(argc, argv);
	Gpu gpu; // context comes up while the inputs are parsed
	const double frag_mean = cmd.Double('u', 0.0), frag_sd = cmd.Double('s', 0.0);
	const int min_read = cmd.Int('n'), max_read = cmd.Int('x');
	const int query = cmd.IsSet('i') ? cmd.Int('i') : -1;

	std::map<int, std::vector<Location>> regions;
	ReadRegionPairs(cmd.Str('r'), regions);
	if (query >= 0)
	{
		auto it = regions.find(query);
		if (it == regions.end())
		{
			std::cerr << "Error: Unable to find fusion " << query << std::endl;
			ExitNow(1);
		}
		std::map<int, std::vector<Location>> only;
		only[query] = it->second;
		regions.swap(only);
	}
	FastaIndex reference;
	reference.Open(cmd.Str('f'));
	ExonModel exons;
	{
		std::ifstream in(cmd.Str('e').c_str());
		if (!in.good() || !exons.Read(in))
		{
			std::cerr << "Error: Unable to read exon regions file " << cmd.Str('e') << std::endl;
			ExitNow(1);
		}
	}
	std::unordered_map<int, ClusterTask> tasks;
	for (const auto& kv : regions)
		InitializeTask(tasks[kv.first], kv.first, kv.second, reference, exons, frag_mean, frag_sd, min_read, max_read);

	ReadIndexFile read_index;
	read_index.Open(cmd.Str('p'));
	std::ifstream align_file(cmd.Str('a').c_str());
	if (!align_file.good())
	{
		std::cerr << "Error: Unable to open " << cmd.Str('a') << std::endl;
		ExitNow(1);
	}

	// ---- every record of a fusion we hold = one task (runs of equal fusion id, SplitAlignment.cpp:319-370) ----
	// The reference handles one run at a time and prints it before it reads the next (tools/splitseq.cpp:104-124): what it
	// dies on -- a malformed line, met while the run in front of it is still being read, or a record whose alignment
	// does not exist -- comes after the text of the runs in front.  `fault` holds that message until they are printed.
	std::string fault;
	std::vector<Record> records;
	std::vector<size_t> run_begin; // first record of each printed run
	{
		std::string line;
		std::vector<std::string> f;
		int run_id = 0;
		bool in_run = false, keep = false;
		while (std::getline(align_file, line))
		{
			SplitChar(line, '\t', f);
			// the run being read is lost with a line that cannot be read: the reference meets such a line while it still
			// collects that run (it looks one line ahead, SplitAlignment.cpp:319-370)
			auto drop_current_run = [&]() {
				if (in_run && keep)
				{
					records.resize(run_begin.back());
					run_begin.pop_back();
				}
			};
			Record r;
			if (f.size() < 9)
			{
				fault = "Error: Format error for candidate reads line:\n" + line + "\n";
				drop_current_run();
				break;
			}
			if (!ParseInt(f[0], r.fusion_id))
			{
				fault = "Error: bad lexical cast: fusion id '" + f[0] + "'\n";
				drop_current_run();
				break;
			}
			// the other fields are only read once the line belongs to the run being collected: a bad one on the first line
			// of a new run comes after the run in front of it has been handled
			int unused = 0;
			const char* bad = nullptr;
			size_t bad_field = 0;
			if (!ParseInt(f[1], r.fragment)) bad = "fragment index", bad_field = 1;
			else if (!ParseInt(f[2], r.read_end)) bad = "read end", bad_field = 2;
			else if (f[3] != "0" && f[3] != "1") bad = "revComp", bad_field = 3;
			else if (!ParseInt(f[4], r.ref_split.first)) bad = "ref split", bad_field = 4;
			else if (!ParseInt(f[5], r.ref_split.second)) bad = "ref split", bad_field = 5;
			else if (!ParseInt(f[6], unused)) bad = "read split", bad_field = 6;
			else if (!ParseInt(f[7], unused)) bad = "read split", bad_field = 7;
			else if (!ParseInt(f[8], unused)) bad = "score", bad_field = 8;
			if (bad)
			{
				fault = std::string("Error: bad lexical cast: ") + bad + " '" + f[bad_field] + "'\n";
				if (in_run && r.fusion_id == run_id) drop_current_run();
				break;
			}
			r.rev_comp = f[3][0] - '0';
			if (!in_run || r.fusion_id != run_id)
			{
				in_run = true;
				run_id = r.fusion_id;
				keep = tasks.find(run_id) != tasks.end(); // runs of other fusions are skipped (splitseq.cpp:113-116)
				if (keep) run_begin.push_back(records.size());
			}
			if (keep) records.push_back(r);
		}
	}
	run_begin.push_back(records.size());

	// ---- tables: window pairs of the fusions seen, reads in the orientation the record names ----
	TableBuilder windows, reads;
	std::unordered_map<int, int> slot;
	int64_t n = (int64_t)records.size();
	std::vector<int32_t> task_cluster((size_t)n), task_read((size_t)n), task_min_score((size_t)n), best((size_t)n);
	std::string seq;
	for (int64_t t = 0; t < n; t++)
	{
		const Record& r = records[(size_t)t];
		auto ins = slot.emplace(r.fusion_id, (int)(windows.Count() / 2));
		if (ins.second)
		{
			const ClusterTask& task = tasks[r.fusion_id];
			windows.Add(task.window[0]);
			windows.Add(task.window[1]);
		}
		read_index.Find(r.fragment, r.read_end, seq);
		if (r.rev_comp) ReverseComplementInPlace(seq);
		task_cluster[(size_t)t] = ins.first->second;
		task_read[(size_t)t] = (int32_t)reads.Add(seq);
		task_min_score[(size_t)t] = (int)((float)seq.length() * (float)kMatch * 0.90); // SplitAlignment.cpp:379
	}

	// ---- GPU: which alignments exist; then the match lists of the one each record names ----
	const dfb_split_params params{kMatch, kMismatch, kGap, 0, kMinAnchor * kMatch};
	const dfb_seq_table window_table = windows.View(), read_table = reads.View();
	std::vector<int32_t> split1((size_t)n), split2((size_t)n), read_split((size_t)n);
	std::vector<int64_t> match_off((size_t)(2 * n + 1), 0);
	std::vector<int32_t> matches;
	if (n > 0)
	{
		if (dfb_split_align_batch(gpu.ctx(), &params, &window_table, &read_table, task_cluster.data(), task_read.data(),
		                          task_min_score.data(), n, best.data()) != DFB_OK)
			gpu.Die("split alignment failed");
		const dfb_split_row* rows = nullptr;
		const int32_t* cols = nullptr;
		int64_t n_rows = 0, n_cols = 0;
		if (dfb_split_result_view(gpu.ctx(), &rows, &n_rows, &cols, &n_cols) != DFB_OK) gpu.Die("split alignment failed");
		int64_t cursor = 0;
		for (int64_t t = 0; t < n; t++)
		{
			const Record& r = records[(size_t)t];
			const int64_t ref2_len = windows.off[(size_t)(2 * task_cluster[(size_t)t] + 2)] - windows.off[(size_t)(2 * task_cluster[(size_t)t] + 1)];
			bool found = false;
			// rows ascending in read_split, columns ascending: the first hit is the tuple Align keeps for this
			// refSplit (SplitAlignment.cpp:381-392) and ReAlign returns (:455-461)
			for (; cursor < n_rows && rows[cursor].task == t; cursor++)
			{
				if (found) continue;
				const dfb_split_row& row = rows[cursor];
				const int32_t* c1 = cols + row.col_begin;
				const int32_t* c2 = c1 + row.n1;
				const bool has1 = std::binary_search(c1, c1 + row.n1, r.ref_split.first);
				const bool has2 = std::binary_search(c2, c2 + row.n2, (int32_t)(ref2_len - r.ref_split.second - 1));
				if (has1 && has2)
				{
					found = true;
					read_split[(size_t)t] = row.read_split;
				}
			}
			if (!found)
			{
				// DebugCheck(false) in ReAlign (SplitAlignment.cpp:462): the runs in front of this record's are still printed
				std::ostringstream msg;
				msg << "Error: false failed: no alignment of read " << r.fragment << (r.read_end == 0 ? "/1" : "/2") << " to fusion "
				    << r.fusion_id << " has ref split " << r.ref_split.first << "," << r.ref_split.second << std::endl;
				fault = msg.str();
				size_t run = 0;
				while (run + 1 < run_begin.size() && run_begin[run + 1] <= (size_t)t) run++;
				n = (int64_t)run_begin[run];
				run_begin.resize(run + 1);
				break;
			}
			split1[(size_t)t] = r.ref_split.first;
			split2[(size_t)t] = r.ref_split.second;
		}
		const int64_t cap = reads.off.back();
		matches.resize((size_t)(2 * cap + 2));
		int64_t n_pairs = 0;
		if (n > 0 && dfb_split_backtrace_batch(gpu.ctx(), &params, &window_table, &read_table, task_cluster.data(), task_read.data(),
		                                       split1.data(), split2.data(), read_split.data(), n, match_off.data(), matches.data(),
		                                       cap, &n_pairs) != DFB_OK)
			gpu.Die("backtrace failed");
	}

	// ---- SplitAlignmentTask::WriteAlignText (SplitAlignment.cpp:626-635) over the text of :402-438 ----
	std::ostringstream os;
	for (size_t run = 0; run + 1 < run_begin.size(); run++)
	{
		const size_t a = run_begin[run], b = run_begin[run + 1];
		if (a == b) continue;
		const ClusterTask& task = tasks[records[a].fusion_id];
		os << task.fusion_id << "\n" << task.window[0] << "|" << task.window[1] << "\n";
		for (size_t t = a; t < b; t++)
		{
			const Record& r = records[t];
			const char* read = reads.bytes.data() + reads.off[(size_t)task_read[t]];
			os << r.fragment << (r.read_end == 0 ? "/1" : "/2") << "\n";
			int prev = -1;
			for (int64_t k = match_off[2 * t]; k < match_off[2 * t + 1]; k++)
			{
				const int ref_pos = matches[(size_t)(2 * k)], read_pos = matches[(size_t)(2 * k + 1)];
				os << std::string((size_t)(ref_pos - prev - 1), prev == -1 ? ' ' : '-') << read[read_pos];
				prev = ref_pos;
			}
			os << std::string((size_t)((int)task.window[0].length() - prev - 1 + 1), '-');
			prev = -1;
			for (int64_t k = match_off[2 * t + 1]; k < match_off[2 * t + 2]; k++)
			{
				const int ref_pos = matches[(size_t)(2 * k)], read_pos = matches[(size_t)(2 * k + 1)];
				os << std::string((size_t)(ref_pos - prev - 1), '-') << read[read_pos];
				prev = ref_pos;
			}
			os << "\n";
		}
	}
	std::cout << os.str();
	if (!fault.empty())
	{
		std::cout.flush();
		std::cerr << fault;
		ExitNow(1);
	}
	FinishProcess(0);
}
