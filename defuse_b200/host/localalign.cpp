// localalign -- drop-in for the reference tool of the same name (tools/localalign.cpp):
//   localalign -m <match> -x <mismatch> -g <gap> [-t <threshold>]  < "id \t reference \t sequence" lines  > "id \t score \t percent"
// Same flags, same input checks and messages, same output bytes and order; SimpleAligner::Align
// (tools/localalign.cpp:79) runs on the GPU in batches through dfb_simple_align_batch.
#include "host_common.h"

#include <string_view>
#include <unordered_map>

using namespace dfbhost;

namespace
{
struct Pending
{
	std::string id;
	int seq_len;
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
	const dfb_simple_params params{match, mismatch, gap};

	// one batch: distinct references (the pipeline repeats each 2001-bp reference on ~100 lines,
	// scripts/prep_local_alignment_seqs.pl:120), sequences, tasks in input order
	TableBuilder refs, seqs;
	std::unordered_map<std::string, int32_t> ref_index;
	std::vector<int32_t> task_ref, task_seq;
	std::vector<Pending> pending;
	std::vector<int32_t> score;
	size_t kBatchTasks = 1u << 20;
	if (const char* e = getenv("DFB_TOOL_BATCH")) kBatchTasks = (size_t)std::max(1, atoi(e)); // tests: force several batches
	const size_t kBatchBytes = 1u << 28;

	std::ios::sync_with_stdio(false);
	std::string out_buf;
	out_buf.reserve(1 << 22);

	auto flush = [&]() {
		if (pending.empty()) return;
		score.resize(pending.size());
		dfb_seq_table rt = refs.View(), st = seqs.View();
		if (dfb_simple_align_batch(gpu.ctx(), &params, &rt, &st, task_ref.data(), task_seq.data(), (int64_t)pending.size(),
		                           score.data()) != DFB_OK)
			gpu.Die("alignment failed");
		std::ostringstream os;
		for (size_t k = 0; k < pending.size(); k++)
		{
			const int max_score = pending[k].seq_len * match;           // tools/localalign.cpp:81
			const double percent = (double)score[k] / (double)max_score; // :82  (0/0 prints as -nan, like the reference)
			if (percent < threshold) continue;                           // :84
			os << pending[k].id << "\t" << score[k] << "\t" << percent << "\n";
		}
		std::cout << os.str();
		std::cout.flush();
		refs.Clear();
		seqs.Clear();
		ref_index.clear();
		task_ref.clear();
		task_seq.clear();
		pending.clear();
	};

	std::string line;
	int line_number = 0;
	std::vector<std::string> fields;
	while (std::getline(std::cin, line))
	{
		line_number++;
		if (line.length() == 0)
		{
			flush();
			std::cerr << "Error: Empty line " << line_number << std::endl;
			exit(1);
		}
		SplitChar(line, '\t', fields);
		if (fields.size() < 3)
		{
			flush();
			std::cerr << "Error: Format error for line " << line_number << std::endl;
			exit(1);
		}
		auto it = ref_index.find(fields[1]);
		int32_t r;
		if (it == ref_index.end())
		{
			r = (int32_t)refs.Add(fields[1]);
			ref_index.emplace(fields[1], r);
		}
		else
		{
			r = it->second;
		}
		task_ref.push_back(r);
		task_seq.push_back((int32_t)seqs.Add(fields[2]));
		pending.push_back(Pending{fields[0], (int)fields[2].size()});
		if (pending.size() >= kBatchTasks || refs.bytes.size() + seqs.bytes.size() >= kBatchBytes) flush();
	}
	flush();
	FinishProcess(0);
}
