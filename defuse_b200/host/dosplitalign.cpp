// dosplitalign -- drop-in for the reference tool of the same name (tools/dosplitalign.cpp):
//   dosplitalign -f ref.fa -e exons.regions -u <frag mean> -s <frag sd> -n <minread> -x <maxread>
//                -r clusters.regions -i improper.sam -1 reads.1.fastq -2 reads.2.fastq -a out.alignments
// Same flags, inputs, candidate enumeration ORDER and output bytes as the reference built against
// libstdc++: the containers whose iteration order decides the record order
// (unordered_map<int,task>, unordered_set<string> of transcripts, unordered_set<int> of overlapping
// cluster ends) are the same std containers filled by the same sequence of inserts.
// SplitReadAligner::Align + GetAlignments (tools/SplitAlignment.cpp:376-379) run on the GPU through
// dfb_split_align_batch; de-duplication and min(score1,score2) (:381-400) stay on the host.
#include "split_tasks.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <memory>
#include <thread>
#include <unordered_map>
#include <unordered_set>

using namespace dfbhost;

namespace
{
struct Candidate
{
	int cluster_id; // fusion id
	int read_id;    // PackId(fragment, end) of the read to align (the OTHER end of the SAM record's fragment)
	int rev_comp;
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
	    {'i', "improper", "Improper Alignments Sam Filename", true, "string", "", false},
	    {'1', "seq1", "End 1 Sequences", true, "string", "", false},
	    {'2', "seq2", "End 2 Sequences", true, "string", "", false},
	    {'a', "align", "Split Alignments Filename", true, "string", "", false},
	});
	cmd.Parse(argc, argv);
	PhaseTimer timer;
	// one context per GPU (DFB_DEVICES), created in the background while the inputs are parsed; candidates are
	// dealt out by cluster, no exchange between GPUs
	std::vector<std::unique_ptr<Gpu>> gpus;
	for (int dev : DeviceList()) gpus.emplace_back(new Gpu(dev));
	const double frag_mean = cmd.Double('u', 0.0), frag_sd = cmd.Double('s', 0.0);
	const int min_read = cmd.Int('n'), max_read = cmd.Int('x');

	// ---- clusters -> tasks (tools/dosplitalign.cpp:80-91, SplitAlignment.cpp:657-686) ----
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
			exit(1);
		}
	}
	std::unordered_map<int, ClusterTask> tasks; // its iteration order is the order mate regions are registered in
	for (const auto& kv : regions)
		InitializeTask(tasks[kv.first], kv.first, kv.second, reference, exons, frag_mean, frag_sd, min_read, max_read);

	timer.Lap("regions, fasta, exons, tasks");
	BinnedRegions binned(2000);
	std::unordered_map<int, int> cluster_slot; // fusion id -> dense index (window pair 2k, 2k+1)
	TableBuilder windows;
	for (const auto& kv : tasks)
	{
		const ClusterTask& t = kv.second;
		cluster_slot[t.fusion_id] = (int)(windows.Count() / 2);
		windows.Add(t.window[0]);
		windows.Add(t.window[1]);
		for (int end = 0; end <= 1; end++)
			for (const Location& loc : t.mate_regions[end]) binned.Add(PackId(t.fusion_id, end), loc);
	}

	FastqReader streams[2];
	const bool ok0 = streams[0].Open(cmd.Str('1'));
	const bool ok1 = streams[1].Open(cmd.Str('2'));
	if (!ok0 || !ok1)
	{
		std::cout << "Error: unable to read sequences" << std::endl;
		exit(1);
	}

	timer.Lap("bins");
	const int n_gpus = (int)gpus.size();

	timer.Lap("gpu contexts");
	// ---- candidates, in the reference's order (SplitAlignment.cpp:266-303): SAM record order x iteration order of
	//      the overlap set; once per (cluster, read id, revComp) ----
	std::vector<Candidate> candidates;
	std::unordered_set<int> needed_reads;
	{
		std::istream* in = &std::cin;
		std::ifstream file;
		if (cmd.Str('i') != "-")
		{
			file.open(cmd.Str('i').c_str());
			if (!file.good())
			{
				std::cerr << "Error: Unable to open sam file " << cmd.Str('i') << std::endl;
				exit(1);
			}
			in = &file;
		}
		std::unordered_map<int, std::unordered_set<std::pair<int, int>, PairHash>> seen;
		std::string line;
		std::vector<std::string> f, q;
		int line_number = 0;
		while (std::getline(*in, line))
		{
			line_number++;
			if (line.empty())
			{
				std::cerr << "Error: Empty alignment line " << line_number << std::endl;
				exit(1);
			}
			if (line[0] == '@') continue;
			SplitChar(line, '\t', f);
			if (f.size() < 10)
			{
				std::cerr << "Error: Format error for alignment line " << line_number << std::endl;
				exit(1);
			}
			const int flag = IntOrDie(f[1], "flag");
			const int pos = IntOrDie(f[3], "pos");
			if (f[2] == "*") continue;
			const int strand = (flag & 0x0010) == 0 ? kPlus : kMinus;
			std::string fragment;
			int read_end = 0;
			SplitChar(f[0], '/', q);
			if (q.size() == 2)
			{
				if (q[1] != "1" && q[1] != "2")
				{
					std::cerr << "Error: Unable to interpret qname for alignment line " << line_number << std::endl;
					exit(1);
				}
				fragment = q[0];
				read_end = q[1] == "1" ? 0 : 1;
			}
			else
			{
				fragment = f[0];
				if (flag & 0x0040) read_end = 0;
				else if (flag & 0x0080) read_end = 1;
				// (neither flag: the reference leaves readEnd uninitialised, AlignmentStream.cpp:108-116; we use end 1)
			}
			const Region r{pos, pos + (int)f[9].length() - 1};
			std::unordered_set<int> overlapping;
			binned.Overlapping(f[2], strand, r, overlapping);
			if (overlapping.empty()) continue;
			const int fragment_index = IntOrDie(fragment, "fragment index");
			for (int cluster_end_id : overlapping)
			{
				const int cluster_id = IdIndex(cluster_end_id), cluster_end = IdEnd(cluster_end_id);
				const int read_id = PackId(fragment_index, read_end == 0 ? 1 : 0);
				const int rev_comp = cluster_end == 0 ? 1 : 0;
				if (seen[cluster_id].insert(std::make_pair(read_id, rev_comp)).second)
				{
					candidates.push_back(Candidate{cluster_id, read_id, rev_comp});
					needed_reads.insert(read_id);
				}
			}
		}
	}

	timer.Lap("sam -> candidates");
	// ---- the reads the candidates need (the reference keeps every read of both files, SplitAlignment.cpp:253-264;
	//      a read id that is absent aligns as the empty string, :286) ----
	std::unordered_map<int, std::string> reads;
	for (int file = 0; file <= 1; file++)
	{
		FastqRead rd;
		while (streams[file].Next(rd))
		{
			const int id = PackId(IntOrDie(rd.fragment, "fragment index"), rd.read_end);
			if (needed_reads.find(id) != needed_reads.end()) reads[id] = rd.sequence;
		}
	}

	timer.Lap("fastq");
	std::ofstream out(cmd.Str('a').c_str());
	if (!out.good())
	{
		std::cerr << "Error: Unable to open " << cmd.Str('a') << std::endl;
		exit(1);
	}

	// ---- align in batches, write records in candidate order ----
	const dfb_split_params params{kMatch, kMismatch, kGap, 0, kMinAnchor * kMatch};
	const dfb_seq_table window_table = windows.View();
	size_t kBatch = (size_t)(1u << 20) * (size_t)n_gpus;
	if (const char* e = getenv("DFB_TOOL_BATCH")) kBatch = (size_t)std::max(1, atoi(e)); // tests: force several batches
	struct Shard
	{
		TableBuilder reads;
		std::vector<int32_t> task_cluster, task_read, task_min_score, best, read_len, ref2_len;
		const dfb_split_row* rows = nullptr;
		const int32_t* cols = nullptr;
		int64_t n_rows = 0, n_cols = 0, cursor = 0;
		int rc = DFB_OK;
	};
	std::vector<Shard> shards((size_t)n_gpus);
	std::vector<int> gpu_of;
	std::string seq;
	for (size_t first = 0; first < candidates.size(); first += kBatch)
	{
		const size_t last = std::min(candidates.size(), first + kBatch);
		const size_t n = last - first;
		// partition by cluster, heaviest first onto the lightest GPU (cost = DP cells); a cluster heavier than a
		// quarter of the mean load is cut into runs of candidates (every GPU holds every window pair)
		gpu_of.assign(n, 0);
		if (n_gpus > 1)
		{
			struct Unit
			{
				double cost;
				std::vector<int32_t> members;
			};
			std::unordered_map<int, std::vector<int32_t>> by_cluster;
			std::vector<int> cluster_order;
			std::vector<double> cost(n);
			double total = 0;
			for (size_t k = 0; k < n; k++)
			{
				const Candidate& c = candidates[first + k];
				const int slot = cluster_slot[c.cluster_id];
				auto it = reads.find(c.read_id);
				const double L = it == reads.end() ? 0.0 : (double)it->second.size();
				cost[k] = L * (double)(windows.off[2 * slot + 2] - windows.off[2 * slot]);
				total += cost[k];
				auto ins = by_cluster.emplace(c.cluster_id, std::vector<int32_t>());
				if (ins.second) cluster_order.push_back(c.cluster_id);
				ins.first->second.push_back((int32_t)k);
			}
			const double cap = std::max(1.0, total / n_gpus / 4.0);
			std::vector<Unit> units;
			for (int cid : cluster_order)
			{
				Unit u{0.0, {}};
				for (int32_t k : by_cluster[cid])
				{
					if (u.cost > cap)
					{
						units.push_back(std::move(u));
						u = Unit{0.0, {}};
					}
					u.members.push_back(k);
					u.cost += cost[k];
				}
				if (!u.members.empty()) units.push_back(std::move(u));
			}
			std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.cost > b.cost; });
			std::vector<double> load((size_t)n_gpus, 0.0);
			for (const Unit& u : units)
			{
				const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
				load[g] += u.cost;
				for (int32_t k : u.members) gpu_of[k] = g;
			}
		}
		for (Shard& sh : shards)
		{
			sh.reads.Clear();
			sh.task_cluster.clear();
			sh.task_read.clear();
			sh.task_min_score.clear();
			sh.read_len.clear();
			sh.ref2_len.clear();
			sh.cursor = 0;
		}
		for (size_t k = 0; k < n; k++)
		{
			const Candidate& c = candidates[first + k];
			Shard& sh = shards[gpu_of[k]];
			auto it = reads.find(c.read_id);
			if (it != reads.end()) seq = it->second; else seq.clear();
			if (c.rev_comp) ReverseComplementInPlace(seq);
			const int slot = cluster_slot[c.cluster_id];
			sh.task_cluster.push_back(slot);
			sh.task_read.push_back((int32_t)sh.reads.Add(seq));
			sh.task_min_score.push_back((int)((float)seq.length() * (float)kMatch * 0.90)); // SplitAlignment.cpp:379
			sh.read_len.push_back((int32_t)seq.size());
			sh.ref2_len.push_back((int32_t)(windows.off[2 * slot + 2] - windows.off[2 * slot + 1]));
		}
		auto run_shard = [&](int g) {
			Shard& sh = shards[g];
			sh.best.resize(sh.task_cluster.size());
			const dfb_seq_table read_table = sh.reads.View();
			sh.rc = dfb_split_align_batch(gpus[g]->ctx(), &params, &window_table, &read_table, sh.task_cluster.data(),
			                              sh.task_read.data(), sh.task_min_score.data(), (int64_t)sh.task_cluster.size(),
			                              sh.best.data());
			if (sh.rc == DFB_OK) sh.rc = dfb_split_result_view(gpus[g]->ctx(), &sh.rows, &sh.n_rows, &sh.cols, &sh.n_cols);
		};
		if (n_gpus == 1)
		{
			run_shard(0);
		}
		else
		{
			std::vector<std::thread> th;
			for (int g = 0; g < n_gpus; g++) th.emplace_back(run_shard, g);
			for (auto& t : th) t.join();
		}
		for (int g = 0; g < n_gpus; g++)
			if (shards[g].rc != DFB_OK) gpus[g]->Die("split alignment failed");

		// merge: candidates in their original order; every shard's rows are in its own task order
		std::ostringstream os;
		std::unordered_set<std::pair<int, int>, PairHash> ref_splits;
		std::vector<int32_t> local_index((size_t)n_gpus, 0);
		for (size_t k = 0; k < n; k++)
		{
			Shard& sh = shards[gpu_of[k]];
			const int32_t t = local_index[gpu_of[k]]++;
			if (sh.cursor >= sh.n_rows || sh.rows[sh.cursor].task != t) continue;
			const Candidate& c = candidates[first + k];
			ref_splits.clear();
			for (; sh.cursor < sh.n_rows && sh.rows[sh.cursor].task == t; sh.cursor++)
			{
				const dfb_split_row& row = sh.rows[sh.cursor];
				const int32_t* c1 = sh.cols + row.col_begin;
				const int32_t* c2 = c1 + row.n1;
				const int score = std::min(row.score1, row.score2); // SplitAlignment.cpp:400
				for (int a = 0; a < row.n1; a++)
				{
					for (int b = 0; b < row.n2; b++)
					{
						const std::pair<int, int> ref_split(c1[a], sh.ref2_len[t] - c2[b] - 1); // SplitReadAligner.cpp:277-278
						if (!ref_splits.insert(ref_split).second) continue;                     // first per refSplit (:383-390)
						os << c.cluster_id << "\t" << IdIndex(c.read_id) << "\t" << IdEnd(c.read_id) << "\t" << c.rev_comp << "\t"
						   << ref_split.first << "\t" << ref_split.second << "\t" << row.read_split << "\t"
						   << sh.read_len[t] - row.read_split << "\t" << score << "\t" << "\n";
					}
				}
			}
			if (os.tellp() > (1 << 22))
			{
				out << os.str();
				os.str(std::string());
			}
		}
		out << os.str();
	}
	out.flush();
	out.close();
	timer.Lap("align + write");
	FinishProcess(0);
}
