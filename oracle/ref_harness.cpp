// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// C-callable harness around the UNMODIFIED reference aligner classes, compiled
// from /root/reference/tools where they lie (see oracle/Makefile):
//   SplitReadAligner  (tools/SplitReadAligner.h:32-67, .cpp:19-298)
//   SimpleAligner     (tools/SimpleAligner.h:17-33,   .cpp:18-63)
// It exists so that tests/ and bench.py's cpu_baseline leg can (a) validate the
// plain-C restatement in oracle/dp_oracle.c and (b) time the reference's own CPU
// implementation.  Built only into oracle/_ref/libref_aligners.so.

#include "SplitReadAligner.h"
#include "SimpleAligner.h"

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

// One SimpleAligner::Align call (reference: tools/localalign.cpp:79, tools/matealign.cpp:209).
int ref_simple_align(int match, int mismatch, int gap,
                     const char* ref, int ref_len, const char* seq, int seq_len)
{
	SimpleAligner aligner(match, mismatch, gap);
	return aligner.Align(std::string(ref, ref_len), std::string(seq, seq_len));
}

// Batch of SimpleAligner::Align calls over CSR-packed byte tables; one aligner
// object re-used like the tools do.  Returns the number of tasks processed.
int64_t ref_simple_align_batch(int match, int mismatch, int gap,
                               const uint8_t* ref_bytes, const int64_t* ref_off,
                               const uint8_t* seq_bytes, const int64_t* seq_off,
                               const int32_t* task_ref, const int32_t* task_seq,
                               int64_t n_tasks, int32_t* out_score)
{
	SimpleAligner aligner(match, mismatch, gap);
	for (int64_t t = 0; t < n_tasks; t++)
	{
		int r = task_ref[t];
		int s = task_seq[t];
		std::string ref((const char*)ref_bytes + ref_off[r], (size_t)(ref_off[r + 1] - ref_off[r]));
		std::string seq((const char*)seq_bytes + seq_off[s], (size_t)(seq_off[s + 1] - seq_off[s]));
		out_score[t] = aligner.Align(ref, seq);
	}
	return n_tasks;
}

// One SplitReadAligner::Align + GetAlignments(minScore, forceSplits=true,
// firstOnly=false, backtrace=false) call, exactly as tools/SplitAlignment.cpp:376-379
// drives it.  Each alignment is written as 7 ints:
//   refSplit.first, refSplit.second, readSplit.first, readSplit.second, score, score1, score2
// Returns the number of alignments the reference produced (may exceed cap; only
// the first `cap` are written).
int64_t ref_split_align(int match, int mismatch, int gap, int end_gaps, int min_split_score,
                        const char* read, int read_len,
                        const char* ref1, int ref1_len,
                        const char* ref2, int ref2_len,
                        int min_score, int32_t* out, int64_t cap)
{
	SplitReadAligner aligner(match, mismatch, gap, end_gaps != 0, min_split_score);
	aligner.Align(std::string(read, read_len), std::string(ref1, ref1_len), std::string(ref2, ref2_len));
	SplitReadAlignVec alignments;
	aligner.GetAlignments(alignments, min_score, true, false, false);
	int64_t n = (int64_t)alignments.size();
	for (int64_t k = 0; k < n && k < cap; k++)
	{
		const SplitReadAlignment& a = alignments[k];
		int32_t* o = out + 7 * k;
		o[0] = a.refSplit.first;
		o[1] = a.refSplit.second;
		o[2] = a.readSplit.first;
		o[3] = a.readSplit.second;
		o[4] = a.score;
		o[5] = a.score1;
		o[6] = a.score2;
	}
	return n;
}

// Batch form for timing: tasks are (cluster, read) pairs; cluster c owns
// reference windows 2c (ref1) and 2c+1 (ref2) of the CSR ref table.
// out_count[t] receives the number of alignments of task t; the alignments are
// appended to `out` (7 ints each) while room remains.  Returns total alignments.
int64_t ref_split_align_batch(int match, int mismatch, int gap, int end_gaps, int min_split_score,
                              const uint8_t* ref_bytes, const int64_t* ref_off,
                              const uint8_t* read_bytes, const int64_t* read_off,
                              const int32_t* task_cluster, const int32_t* task_read,
                              const int32_t* task_min_score, int64_t n_tasks,
                              int32_t* out_count, int32_t* out, int64_t cap)
{
	SplitReadAligner aligner(match, mismatch, gap, end_gaps != 0, min_split_score);
	int64_t total = 0;
	for (int64_t t = 0; t < n_tasks; t++)
	{
		int c = task_cluster[t];
		int r = task_read[t];
		std::string ref1((const char*)ref_bytes + ref_off[2 * c], (size_t)(ref_off[2 * c + 1] - ref_off[2 * c]));
		std::string ref2((const char*)ref_bytes + ref_off[2 * c + 1], (size_t)(ref_off[2 * c + 2] - ref_off[2 * c + 1]));
		std::string read((const char*)read_bytes + read_off[r], (size_t)(read_off[r + 1] - read_off[r]));
		aligner.Align(read, ref1, ref2);
		SplitReadAlignVec alignments;
		aligner.GetAlignments(alignments, task_min_score[t], true, false, false);
		out_count[t] = (int32_t)alignments.size();
		for (size_t k = 0; k < alignments.size(); k++)
		{
			if (total < cap)
			{
				const SplitReadAlignment& a = alignments[k];
				int32_t* o = out + 7 * total;
				o[0] = a.refSplit.first;
				o[1] = a.refSplit.second;
				o[2] = a.readSplit.first;
				o[3] = a.readSplit.second;
				o[4] = a.score;
				o[5] = a.score1;
				o[6] = a.score2;
			}
			total++;
		}
	}
	return total;
}

// GetAlignments(minScore, forceSplits=true, firstOnly=false, backtrace=TRUE): the match lists of alignment
// number `which` in emission order (tools/SplitReadAligner.cpp:287-292), as interleaved (refPos, readPos).
// Returns the number of alignments the reference produced, or -1 when `which` is out of range.
int64_t ref_split_backtrace(int match, int mismatch, int gap, int end_gaps, int min_split_score,
                            const char* read, int read_len, const char* ref1, int ref1_len,
                            const char* ref2, int ref2_len, int min_score, int64_t which,
                            int32_t* header /* 7 ints as in ref_split_align */,
                            int32_t* matches1, int32_t* n1, int32_t* matches2, int32_t* n2)
{
	SplitReadAligner aligner(match, mismatch, gap, end_gaps != 0, min_split_score);
	aligner.Align(std::string(read, read_len), std::string(ref1, ref1_len), std::string(ref2, ref2_len));
	SplitReadAlignVec alignments;
	aligner.GetAlignments(alignments, min_score, true, false, true);
	if (which < 0 || which >= (int64_t)alignments.size()) return -1;
	const SplitReadAlignment& a = alignments[which];
	header[0] = a.refSplit.first;
	header[1] = a.refSplit.second;
	header[2] = a.readSplit.first;
	header[3] = a.readSplit.second;
	header[4] = a.score;
	header[5] = a.score1;
	header[6] = a.score2;
	*n1 = (int32_t)a.matches1.size();
	for (size_t k = 0; k < a.matches1.size(); k++)
	{
		matches1[2 * k] = a.matches1[k].first;
		matches1[2 * k + 1] = a.matches1[k].second;
	}
	*n2 = (int32_t)a.matches2.size();
	for (size_t k = 0; k < a.matches2.size(); k++)
	{
		matches2[2 * k] = a.matches2[k].first;
		matches2[2 * k + 1] = a.matches2[k].second;
	}
	return (int64_t)alignments.size();
}

// ReverseComplement of tools/Common.cpp:32-54 (ACGTacgt only), for fixture generation.
void ref_reverse_complement(char* seq, int len)
{
	std::string s(seq, len);
	ReverseComplement(s);
	memcpy(seq, s.data(), len);
}

}  // extern "C"
