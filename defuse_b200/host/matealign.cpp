// matealign -- drop-in for the reference tool of the same name (tools/matealign.cpp):
//   matealign -m -x -g [-t] -s <searchlength> -r <reference.fa> -1 <reads1.fastq> -2 <reads2.fastq>  < SAM  > "fragment \t score \t percent"
// For every read whose OTHER end has alignments in the SAM, the read is aligned against the
// searchlength+1 window downstream of each of those alignments (N-padded at the sequence ends),
// in fastq order x SAM order (tools/matealign.cpp:179-223).  SimpleAligner::Align (:209) runs on
// the GPU in batches; everything else is the same host logic, written for batching.
#include "host_common.h"

#include <fstream>
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
		std::ifstream in(filename.c_str());
		if (!in.good())
		{
			std::cerr << "Error: unable to open file " << filename << std::endl;
			exit(1);
		}
		std::string id, sequence, line;
		while (std::getline(in, line))
		{
			if (line.empty()) continue;
			if (line[0] == '>')
			{
				if (!id.empty()) seqs[id] = sequence;
				id = line.substr(1);
				sequence.clear();
			}
			else
			{
				sequence.append(line);
			}
		}
		if (!id.empty()) seqs[id] = sequence;
	}
	// [start, end] 1-based inclusive, 'N' where the window leaves the sequence (tools/Sequences.cpp:60-79)
	void Window(const std::string& id, int start, int end, std::string& out) const
	{
		auto it = seqs.find(id);
		if (it == seqs.end())
		{
			std::cerr << "Error: Unable to find sequence " << id << std::endl;
			exit(1);
		}
		const std::string& full = it->second;
		const long long len = (long long)full.size();
		const long long seq_start = std::max<long long>(1, start);
		const long long prepend = seq_start - start;
		const long long seq_end = std::min<long long>(len, end);
		const long long append = (long long)end - seq_end;
		out.assign((size_t)prepend, 'N');
		if (seq_start - 1 > len || append < 0)
		{
			// std::string::substr / string(n,'N') would throw here in the reference (window starts beyond the
			// sequence, or ends before position 1): it aborts; we report and fail the same way (non-zero exit)
			std::cerr << "Error: window " << start << "-" << end << " outside sequence " << id << std::endl;
			exit(1);
		}
		const long long take = seq_end - seq_start + 1;
		if (take > 0) out.append(full, (size_t)(seq_start - 1), (size_t)take);
		else if (take < 0) out.append(full, (size_t)(seq_start - 1), std::string::npos); // substr(pos, huge) semantics
		out.append((size_t)append, 'N');
	}
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
	const dfb_simple_params params{match, mismatch, gap};

	// ---- SAM on stdin -> alignments per read id, in input order (tools/matealign.cpp:80-158) ----
	std::unordered_map<int, std::vector<MatePosition>> read_alignments;
	std::unordered_map<std::string, int> ref_lookup;
	std::vector<std::string> ref_names;
	std::ios::sync_with_stdio(false);
	std::string line;
	int line_number = 0;
	std::vector<std::string> f, q;
	while (std::getline(std::cin, line))
	{
		line_number++;
		if (line.length() == 0)
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
		const int strand = (flag & 0x0010) == 0 ? 0 : 1;
		SplitChar(f[0], '/', q);
		if (q.size() != 2 || (q[1] != "1" && q[1] != "2"))
		{
			std::cerr << "Error: Unable to interpret qname for alignment line " << line_number << std::endl;
			exit(1);
		}
		const int read_end = q[1] == "1" ? 0 : 1;
		const int start = pos;
		const int end = start + (int)f[9].length() - 1;
		const int fragment_index = IntOrDie(q[0], "fragment index");
		auto ins = ref_lookup.emplace(f[2], (int)ref_names.size());
		if (ins.second) ref_names.push_back(f[2]);
		MatePosition mp;
		mp.ref_index = ins.first->second;
		mp.strand = strand;
		mp.position = strand == 0 ? start : end;
		read_alignments[PackId(fragment_index, read_end)].push_back(mp);
	}
	std::cerr << "Read alignments" << std::endl;

	FastaSequences reference;
	reference.Read(reference_fasta);
	std::cerr << "Read reference fasta" << std::endl;

	FastqReader streams[2];
	const bool ok0 = streams[0].Open(reads_filename[0]);
	const bool ok1 = streams[1].Open(reads_filename[1]);
	if (!ok0 || !ok1)
	{
		std::cout << "Error: unable to read sequences" << std::endl;
		exit(1);
	}

	// ---- tasks: (window, read) in fastq order x SAM order; flushed in batches ----
	TableBuilder windows, reads;
	std::vector<int32_t> task_ref, task_seq, task_fragment, task_len, score;
	size_t kBatchTasks = 1u << 19;
	if (const char* e = getenv("DFB_TOOL_BATCH")) kBatchTasks = (size_t)std::max(1, atoi(e)); // tests: force several batches
	const size_t kBatchBytes = 1u << 28;
	auto flush = [&]() {
		if (task_ref.empty()) return;
		score.resize(task_ref.size());
		dfb_seq_table wt = windows.View(), rt = reads.View();
		if (dfb_simple_align_batch(gpu.ctx(), &params, &wt, &rt, task_ref.data(), task_seq.data(), (int64_t)task_ref.size(),
		                           score.data()) != DFB_OK)
			gpu.Die("alignment failed");
		std::ostringstream os;
		for (size_t k = 0; k < task_ref.size(); k++)
		{
			const int max_score = task_len[k] * match;                   // tools/matealign.cpp:211
			const double percent = (double)score[k] / (double)max_score; // :212
			if (percent < threshold) continue;
			os << task_fragment[k] << "\t" << score[k] << "\t" << percent << "\n";
		}
		std::cout << os.str();
		std::cout.flush();
		windows.Clear();
		reads.Clear();
		task_ref.clear();
		task_seq.clear();
		task_fragment.clear();
		task_len.clear();
	};

	std::string window;
	for (int file = 0; file <= 1; file++)
	{
		FastqRead read;
		while (streams[file].Next(read))
		{
			const int fragment_index = IntOrDie(read.fragment, "fragment index");
			const int other_id = PackId(fragment_index, 1 - read.read_end);
			auto it = read_alignments.find(other_id);
			if (it == read_alignments.end()) continue;
			int32_t read_slot = -1;
			for (const MatePosition& mp : it->second)
			{
				if (mp.strand == 0)
				{
					reference.Window(ref_names[mp.ref_index], mp.position, mp.position + search_length, window);
					ReverseComplementInPlace(window);
				}
				else
				{
					reference.Window(ref_names[mp.ref_index], mp.position - search_length, mp.position, window);
				}
				if (read_slot < 0) read_slot = (int32_t)reads.Add(read.sequence);
				task_ref.push_back((int32_t)windows.Add(window));
				task_seq.push_back(read_slot);
				task_fragment.push_back((int32_t)((unsigned)fragment_index & 0x7fffffffu)); // readID.fragmentIndex is a 31-bit field
				task_len.push_back((int)read.sequence.size());
			}
			if (task_ref.size() >= kBatchTasks || windows.bytes.size() + reads.bytes.size() >= kBatchBytes) flush();
		}
	}
	flush();
	FinishProcess(0);
}
