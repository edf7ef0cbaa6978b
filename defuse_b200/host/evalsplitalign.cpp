// evalsplitalign -- drop-in for the reference tool of the same name (tools/evalsplitalign.cpp): the consumer of
// dosplitalign's sorted records.  Per fusion: sum the record scores per refSplit, take the best split, predict the
// fusion sequence and the two break positions, and pass the supporting records on.
//   evalsplitalign -f ref.fa -e exons.regions -u <frag mean> -s <frag sd> -n <minread> -x <maxread>
//                  -r clusters.regions -a sorted.alignments -q out.seq -b out.break -p out.predalign
// No DP happens here; the tool is host-only and needs no GPU.  Same flags, inputs and output
// bytes as the reference built against libstdc++: the split with the highest summed score is the first such split
// in the iteration order of an unordered_map<pair<int,int>,int> filled in record order (SplitAlignment.cpp:505-528),
// which is reproduced with the same std container and the same hash.
#include "split_tasks.h"

#include <cmath>

using namespace dfbhost;

namespace
{
// one line of dosplitalign output (SplitAlignment::WriteAlignment, tools/SplitAlignment.cpp:305-317)
struct Record
{
	int fusion_id, fragment, read_end, rev_comp;
	std::pair<int, int> ref_split, read_split;
	int score;
};

// lexical_cast<bool>: exactly "0" or "1"
bool ParseBool(const std::string& s, int& out)
{
	if (s == "0" || s == "1")
	{
		out = s[0] - '0';
		return true;
	}
	return false;
}

// SplitAlignment::ReadSortedAlignments (tools/SplitAlignment.cpp:319-370): the run of lines with the first line's
// fusion id.  `pending` carries the first line of the next run (the reference seeks back instead).
bool ReadRun(std::istream& in, std::string& pending, bool& have_pending, std::vector<Record>& run)
{
	run.clear();
	std::string line;
	std::vector<std::string> f;
	int first_id = -1;
	for (;;)
	{
		if (have_pending)
		{
			line.swap(pending);
			have_pending = false;
		}
		else if (!std::getline(in, line))
			break;
		SplitChar(line, '\t', f);
		// the reference checks for 7 fields and then reads fields 7 and 8 (:331,359-360); fewer than 9 is an error here
		if (f.size() < 9)
		{
			std::cerr << "Error: Format error for candidate reads line:" << std::endl << line << std::endl;
			ExitNow(1);
		}
		Record r;
		r.fusion_id = IntOrDie(f[0], "fusion id");
		if (run.empty())
			first_id = r.fusion_id;
		else if (r.fusion_id != first_id)
		{
			pending.swap(line);
			have_pending = true;
			break;
		}
		r.fragment = IntOrDie(f[1], "fragment index");
		r.read_end = IntOrDie(f[2], "read end");
		if (!ParseBool(f[3], r.rev_comp))
		{
			std::cerr << "Error: bad lexical cast: revComp '" << f[3] << "'" << std::endl;
			ExitNow(1);
		}
		r.ref_split = std::make_pair(IntOrDie(f[4], "ref split"), IntOrDie(f[5], "ref split"));
		r.read_split = std::make_pair(IntOrDie(f[6], "read split"), IntOrDie(f[7], "read split"));
		r.score = IntOrDie(f[8], "score");
		run.push_back(r);
	}
	return !run.empty();
}
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
	    {'a', "align", "Split Alignments Filename", true, "string", "", false},
	    {'q', "seq", "Sequences Filename", true, "string", "", false},
	    {'b', "break", "Break Positions Filename", true, "string", "", false},
	    {'p', "predalign", "Prediction Split Alignments Filename", true, "string", "", false},
	});
	cmd.Parse(argc, argv);
	const double frag_mean = cmd.Double('u', 0.0), frag_sd = cmd.Double('s', 0.0);
	const int min_read = cmd.Int('n'), max_read = cmd.Int('x');

	// clusters -> tasks (tools/evalsplitalign.cpp:80-84, SplitAlignment.cpp:657-686)
	std::map<int, std::vector<Location>> regions;
	ReadRegionPairs(cmd.Str('r'), regions);
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

	std::ifstream align_file(cmd.Str('a').c_str());
	std::ofstream seq_file(cmd.Str('q').c_str());
	std::ofstream break_file(cmd.Str('b').c_str());
	std::ofstream pred_file(cmd.Str('p').c_str());
	const struct
	{
		bool ok;
		const std::string& name;
	} files[] = {{align_file.good(), cmd.Str('a')}, {seq_file.good(), cmd.Str('q')}, {break_file.good(), cmd.Str('b')},
	             {pred_file.good(), cmd.Str('p')}};
	for (const auto& f : files)
		if (!f.ok)
		{
			std::cerr << "Error: Unable to open " << f.name << std::endl;
			ExitNow(1);
		}

	std::vector<Record> run, support;
	std::string pending;
	bool have_pending = false;
	while (ReadRun(align_file, pending, have_pending, run))
	{
		const int fusion_id = run.front().fusion_id;
		auto it = tasks.find(fusion_id);
		if (it == tasks.end())
		{
			// the reference default-constructs a task here (operator[]) and trips a DebugCheck on its empty windows
			std::cerr << "Error: no fusion regions for fusion " << fusion_id << std::endl;
			ExitNow(1);
		}
		const ClusterTask& task = it->second;

		// ---- SplitAlignmentTask::Evaluate (tools/SplitAlignment.cpp:484-594) ----
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
		std::string sequence = "N";
		int count = 0;
		double pos_avg = -1.0, min_avg = -1.0;
		int break_pos[2] = {0, 0};
		support.clear();
		if (max_score == -1)
		{
			// (every summed score below 0: the reference reports it and writes the empty prediction, break
			// positions uninitialised there, 0 here)
			std::cerr << "Error: Unable to find max score split" << std::endl;
		}
		else
		{
			for (const Record& r : run)
				if (r.ref_split == best) support.push_back(r);
			if (!(best.first >= 0 && (size_t)best.first <= task.window[0].length()) ||
			    !(best.second + 1 >= 0 && (size_t)(best.second + 1) < task.window[1].length()))
			{
				std::cerr << "Error: split outside the breakpoint windows of fusion " << fusion_id << std::endl;
				ExitNow(1);
			}
			sequence = task.remainder[0] + task.window[0].substr(0, (size_t)best.first) + "|" +
			           task.window[1].substr((size_t)(best.second + 1)) + task.remainder[1];
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

		// BreakPrediction::WriteSequence / WriteBreak / WriteAlignments (:596-624)
		seq_file << fusion_id << "\t" << sequence << "\t" << "0" << "\t" << count << "\t" << pos_avg << "\t" << min_avg << std::endl;
		for (int end = 0; end <= 1; end++)
			break_file << fusion_id << "\t" << end << "\t" << task.align_ref_name[end] << "\t"
			           << (task.align_strand[end] == kPlus ? "+" : "-") << "\t" << break_pos[end] << std::endl;
		for (const Record& r : support)
			pred_file << r.fusion_id << "\t" << r.fragment << "\t" << r.read_end << "\t" << r.rev_comp << "\t" << r.ref_split.first
			          << "\t" << r.ref_split.second << "\t" << r.read_split.first << "\t" << r.read_split.second << "\t" << r.score
			          << "\t" << "\n";
	}
	return 0;
}
