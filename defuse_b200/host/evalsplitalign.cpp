// evalsplitalign -- drop-in for the reference tool of the same name (tools/evalsplitalign.cpp): the consumer of
// dosplitalign's sorted records.  Per fusion: sum the record scores per refSplit, take the best split, predict the
// fusion sequence and the two break positions, and pass the supporting records on.
//   evalsplitalign -f ref.fa -e exons.regions -u <frag mean> -s <frag sd> -n <minread> -x <maxread>
//                  -r clusters.regions -a sorted.alignments -q out.seq -b out.break -p out.predalign
// No DP happens here; the tool is host-only and needs no GPU.  Same flags, inputs and output
// bytes as the reference built against libstdc++: the split with the highest summed score is the first such split
// in the iteration order of an unordered_map<pair<int,int>,int> filled in record order (SplitAlignment.cpp:505-528),
// which is reproduced with the same std container and the same hash.
//
// Ingest: the sorted records file is mapped and cut, block by block, into regions that
// begin where the fusion id changes; every region is parsed in place, evaluated and formatted by one host thread, and
// the regions' outputs are written in file order.  A malformed line still ends the run exactly where a line-by-line
// reader would have ended it (the first one in file order wins, everything in front of it is written).
#include "eval_core.h"

using namespace dfbhost;

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

	MappedInput align_file;
	const bool align_ok = align_file.OpenFile(cmd.Str('a'));
	std::ofstream seq_file(cmd.Str('q').c_str());
	std::ofstream break_file(cmd.Str('b').c_str());
	std::ofstream pred_file(cmd.Str('p').c_str());
	const struct
	{
		bool ok;
		const std::string& name;
	} files[] = {{align_ok, cmd.Str('a')}, {seq_file.good(), cmd.Str('q')}, {break_file.good(), cmd.Str('b')},
	             {pred_file.good(), cmd.Str('p')}};
	for (const auto& f : files)
		if (!f.ok)
		{
			std::cerr << "Error: Unable to open " << f.name << std::endl;
			ExitNow(1);
		}

	if (!EvaluateSortedRecords(align_file.data(), align_file.data() + align_file.size(), tasks, seq_file, break_file, pred_file)) ExitNow(1);
	return 0;
}
