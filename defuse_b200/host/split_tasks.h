// split_tasks.h -- what dosplitalign, evalsplitalign and splitseq share: region/exon/FASTA readers and the
// construction of one SplitAlignmentTask per candidate fusion (tools/SplitAlignment.cpp:31-175, 657-686).
// Host-side only; the DP itself runs on the GPU behind include/defuse_b200.h.
#ifndef DFB_SPLIT_TASKS_H
#define DFB_SPLIT_TASKS_H

#include "host_common.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <memory>
#include <unordered_map>
#include <unordered_set>

namespace dfbhost
{
enum
{
	kPlus = 0,
	kMinus = 1
};

// constants of tools/SplitAlignment.cpp:25-29
const int kMatch = 2, kMismatch = -1, kGap = -2, kMinAnchor = 4;

struct Region
{
	int start, end;
};

struct Location
{
	std::string ref_name;
	int strand = 0, start = 0, end = 0;
};

inline int StrandOrDie(const std::string& s)
{
	if (s == "+") return kPlus;
	if (s == "-") return kMinus;
	std::cerr << "Error: Unable to intepret strand " << s << std::endl;
	ExitNow(1);
}

// clusterID \t clusterEnd \t refName \t +|- \t start \t end   (tools/Parsers.cpp:211-264)
inline void ReadRegionPairs(const std::string& filename, std::map<int, std::vector<Location>>& pairs)
{
	std::ifstream in(filename.c_str());
	if (!in.good())
	{
		std::cerr << "Error: Unable to open align region pairs file " << filename << std::endl;
		ExitNow(1);
	}
	std::string line;
	std::vector<std::string> f;
	while (std::getline(in, line))
	{
		if (line.empty()) continue;
		SplitChar(line, '\t', f);
		if (f.size() < 5) continue;
		int id = 0, end = 0, start = 0, stop = 0;
		auto bad_line = [&]() {
			std::cout << "Failed to interpret region:" << std::endl << line << std::endl;
			ExitNow(1);
		};
		// fields are read in the reference's order (Parsers.cpp:242-251): ids, the pairEnd check, strand, positions --
		// a line with several faults dies on the first of them.  The reference reads field 5 after checking for only 5
		// fields (:237,251): a 5-field line is undefined behaviour there; here it is a format error
		if (f.size() < 6 || !ParseInt(f[0], id) || !ParseInt(f[1], end)) bad_line();
		if (end != 0 && end != 1)
		{
			std::cerr << "Error: pairEnd == 0 || pairEnd == 1 failed for region line: " << line << std::endl;
			ExitNow(1);
		}
		Location loc;
		loc.ref_name = f[2];
		loc.strand = StrandOrDie(f[3]);
		if (!ParseInt(f[4], start) || !ParseInt(f[5], stop)) bad_line();
		loc.start = start;
		loc.end = stop;
		std::vector<Location>& v = pairs[id];
		v.resize(2);
		v[end] = loc;
	}
}

// ---------------------------------------------------------------------------------------------
// exon model: gene \t transcript \t chromosome \t strand \t (start \t end)+   (tools/ExonRegions.cpp:21-112)
// only the queries SplitAlignmentTask::Initialize makes
// ---------------------------------------------------------------------------------------------
class ExonModel
{
public:
	bool Read(std::istream& in)
	{
		std::string line;
		std::vector<std::string> f;
		while (std::getline(in, line))
		{
			if (line.empty()) continue;
			SplitChar(line, '\t', f);
			if (f.size() < 6) continue;
			const std::string& gene = f[0];
			const std::string& transcript = f[1];
			const std::string& chromosome = f[2];
			std::vector<Region> exons;
			for (size_t k = 5; k < f.size(); k += 2)
			{
				Region e;
				if (!ParseInt(f[k - 1], e.start) || !ParseInt(f[k], e.end))
				{
					std::cout << "Failed to interpret exon:" << std::endl << line << std::endl;
					ExitNow(1);
				}
				exons.push_back(e);
			}
			const int strand = StrandOrDie(f[3]);
			int length = 0;
			for (const Region& e : exons) length += e.end - e.start + 1;
			Transcript& t = mTranscripts[transcript];
			t.chromosome = chromosome;
			t.strand = strand;
			t.length = length;
			t.gene = gene;
			t.exons[kPlus] = exons;
			// minus-strand view: negate and reverse, so that "strand space" walks 5'->3' (ExonRegions.cpp:114-123)
			t.exons[kMinus].clear();
			for (auto it = exons.rbegin(); it != exons.rend(); ++it) t.exons[kMinus].push_back(Region{-it->end, -it->start});
			t.span = Region{exons.front().start, exons.back().end};
			const int first_bin = t.span.start / kBinLength, last_bin = t.span.end / kBinLength;
			for (int bin = first_bin; bin <= last_bin; bin++) mLookup[chromosome][bin].push_back(transcript);
		}
		return true;
	}

	bool IsTranscript(const std::string& transcript) const { return mTranscripts.find(transcript) != mTranscripts.end(); }

	const std::string& GeneOf(const std::string& transcript) const
	{
		auto it = mTranscripts.find(transcript);
		if (it == mTranscripts.end())
		{
			std::cerr << "Error: Data mismatch, unable to find gene for transcript " << transcript << std::endl;
			ExitNow(1);
		}
		return it->second.gene;
	}

	// transcripts whose span overlaps the region; ORDER = iteration order of an unordered_set<string>
	// filled bin by bin (ExonRegions.cpp:130-160) -- it decides the order of a cluster's mate regions
	void RegionTranscripts(const std::string& chromosome, const Region& region, std::vector<std::string>& out) const
	{
		auto chr = mLookup.find(chromosome);
		if (chr == mLookup.end())
		{
			std::cerr << "Error: Data mismatch, invalid chromosome " << chromosome << std::endl;
			ExitNow(1);
		}
		std::unordered_set<std::string> unique;
		const int first_bin = region.start / kBinLength, last_bin = region.end / kBinLength;
		for (int bin = first_bin; bin <= last_bin; bin++)
		{
			auto b = chr->second.find(bin);
			if (b == chr->second.end()) continue;
			for (const std::string& name : b->second)
			{
				const Region& span = mTranscripts.find(name)->second.span;
				if (!(span.end < region.start || span.start > region.end)) unique.insert(name);
			}
		}
		out.insert(out.end(), unique.begin(), unique.end());
	}

	// transcript coordinate -> genome coordinate (ExonRegions.cpp:258-304)
	void TranscriptToGenome(const std::string& transcript, int strand, int position, std::string& chromosome, int& out_strand,
	                        int& out_position) const
	{
		const Transcript& t = FindOrDie(transcript);
		const std::vector<Region>& exons = t.exons[kPlus];
		chromosome = t.chromosome;
		out_strand = (t.strand == strand) ? kPlus : kMinus;
		if (t.strand == kMinus) position = t.length - position + 1;
		int offset = 0;
		for (const Region& e : exons)
		{
			const int len = e.end - e.start + 1;
			if (position <= offset + len)
			{
				out_position = position - (offset + 1) + e.start;
				return;
			}
			offset += len;
		}
		out_position = position - t.length + exons.back().end;
	}

	// genome position + extension range -> range in transcript coordinates (ExonRegions.cpp:416-468)
	bool ThroughTranscript(const std::string& transcript, int position, int strand, int extend_min, int extend_max,
	                       int& out_strand, int& start, int& end) const
	{
		const Transcript& t = FindOrDie(transcript);
		const std::vector<Region>& exons = t.exons[strand];
		out_strand = (strand == t.strand) ? kPlus : kMinus;
		const int strand_position = (strand == kPlus) ? position : -position;
		if (strand_position > exons.back().end) return false;
		int offset = 0;
		for (const Region& e : exons)
		{
			if (strand_position <= e.end)
			{
				const int rel_start = strand_position - e.start + extend_min + 1;
				const int rel_end = strand_position - e.start + extend_max + 1;
				if (rel_end < 1) return false;
				start = std::max(1, rel_start) + offset;
				end = std::max(1, rel_end) + offset;
				break;
			}
			offset += e.end - e.start + 1;
		}
		if (end < 1 || start > t.length) return false;
		if (strand != t.strand)
		{
			start = t.length - start + 1;
			end = t.length - end + 1;
			std::swap(start, end);
		}
		return true;
	}

private:
	struct Transcript
	{
		std::string chromosome, gene;
		int strand = 0, length = 0;
		std::vector<Region> exons[2];
		Region span{0, 0};
	};
	const Transcript& FindOrDie(const std::string& transcript) const
	{
		auto it = mTranscripts.find(transcript);
		if (it == mTranscripts.end() || it->second.exons[kPlus].empty())
		{
			std::cerr << "Error: Data mismatch, unable to find transcript " << transcript << std::endl;
			ExitNow(1);
		}
		return it->second;
	}
	static const int kBinLength = 100000;
	std::unordered_map<std::string, Transcript> mTranscripts;
	std::unordered_map<std::string, std::unordered_map<int, std::vector<std::string>>> mLookup;
};

// ---------------------------------------------------------------------------------------------
// faidx random access (external/samtools-0.1.8/faidx.c:305-355 fai_fetch, :62-140 fai_build_core)
// ---------------------------------------------------------------------------------------------
class FastaIndex
{
public:
	void Open(const std::string& fasta)
	{
		mFile = fopen(fasta.c_str(), "rb");
		if (!mFile)
		{
			std::cerr << "[fai_load] fail to open FASTA file." << std::endl;
			ExitNow(1);
		}
		const std::string fai = fasta + ".fai";
		std::ifstream in(fai.c_str());
		if (!in.good())
		{
			Build();
			// like fai_load, leave the index next to the FASTA for the next run (best effort)
			std::ofstream out(fai.c_str());
			if (out.good())
				for (const std::string& name : mOrder)
				{
					const Entry& e = mIndex[name];
					out << name << "\t" << e.len << "\t" << e.offset << "\t" << e.line_blen << "\t" << e.line_len << "\n";
				}
			return;
		}
		std::string line;
		std::vector<std::string> f;
		while (std::getline(in, line))
		{
			SplitChar(line, '\t', f);
			if (f.size() < 5) continue;
			Entry e;
			e.len = atoll(f[1].c_str());
			e.offset = atoll(f[2].c_str());
			e.line_blen = atoi(f[3].c_str());
			e.line_len = atoi(f[4].c_str());
			mIndex[f[0]] = e;
		}
	}

	// FastaIndex::Get (tools/FastaIndex.cpp:22-58): start/length are clamped IN PLACE, the region is fetched
	// 1-based inclusive, clipped at the sequence end; minus strand -> reverse complement
	void Get(const std::string& ref_name, int strand, int& start, int& length, std::string& sequence) const
	{
		if (length < 0)
		{
			sequence.clear();
			return;
		}
		if (start < 1)
		{
			length -= 1 - start;
			start = 1;
		}
		const int end = start + length - 1;
		auto it = mIndex.find(ref_name);
		if (it == mIndex.end())
		{
			std::cerr << "Error: Unable to find sequence for " << ref_name << std::endl;
			ExitNow(1);
		}
		const Entry& e = it->second;
		// fai_fetch's own clamping of "name:start-end" (atoi semantics on the two numbers)
		long long beg = start, stop = end;
		if (beg > 0) --beg;
		if (beg >= e.len) beg = e.len;
		if (stop >= e.len) stop = e.len;
		if (beg > stop) beg = stop;
		sequence.clear();
		if (stop > beg && e.line_blen > 0)
		{
			const long long want = stop - beg;
			sequence.reserve((size_t)want);
			fseeko(mFile, (off_t)(e.offset + beg / e.line_blen * e.line_len + beg % e.line_blen), SEEK_SET);
			int c;
			while ((long long)sequence.size() < want && (c = getc_unlocked(mFile)) != EOF)
				if (c > 32 && c < 127) sequence.push_back((char)c); // isgraph
		}
		length = (int)sequence.size();
		if (strand == kMinus) ReverseComplementInPlace(sequence);
	}

private:
	struct Entry
	{
		long long len = 0, offset = 0;
		int line_blen = 0, line_len = 0;
	};
	void Build()
	{
		// one pass: name = header up to the first whitespace; offset = first base; line lengths from the first line
		fseeko(mFile, 0, SEEK_SET);
		std::string name;
		Entry e;
		bool have = false, first_line = false;
		long long pos = 0, line_start = 0;
		int c;
		std::string line;
		auto close_entry = [&]() {
			if (have)
			{
				if (mIndex.find(name) == mIndex.end()) mOrder.push_back(name);
				mIndex[name] = e;
			}
		};
		while (true)
		{
			line.clear();
			line_start = pos;
			bool eof = true;
			while ((c = getc_unlocked(mFile)) != EOF)
			{
				pos++;
				eof = false;
				if (c == '\n') break;
				line.push_back((char)c);
			}
			if (eof && line.empty()) break;
			const bool had_newline = (c == '\n');
			if (!line.empty() && line[0] == '>')
			{
				close_entry();
				size_t k = 1;
				while (k < line.size() && !isspace((unsigned char)line[k])) k++;
				name = line.substr(1, k - 1);
				e = Entry();
				e.offset = pos;
				have = true;
				first_line = true;
			}
			else if (have)
			{
				int graph = 0;
				for (char ch : line)
					if (ch > 32 && ch < 127) graph++;
				if (first_line && !line.empty())
				{
					e.line_blen = graph;
					e.line_len = (int)line.size() + (had_newline ? 1 : 0);
					first_line = false;
				}
				else if (first_line && line.empty())
				{
					e.offset = pos; // empty line right after the header: the sequence starts behind it
				}
				e.len += graph;
			}
			(void)line_start;
			if (!had_newline) break;
		}
		close_entry();
	}
	FILE* mFile = nullptr;
	std::unordered_map<std::string, Entry> mIndex;
	std::vector<std::string> mOrder;
};

inline bool ParseTranscriptId(const std::string& id, std::string& gene, std::string& transcript)
{
	std::vector<std::string> f;
	SplitChar(id, '|', f);
	if (f.size() < 2) return false;
	gene = f[0];
	transcript = f[1];
	return true;
}

// one candidate fusion: the two breakpoint windows and where the mates of split reads may align
struct ClusterTask
{
	int fusion_id = 0;
	std::string window[2];                 // mSplitAlignSeq
	std::vector<Location> mate_regions[2]; // mMateRegions
	// what Evaluate (break prediction) needs on top (SplitAlignment.cpp:49-103)
	std::string align_ref_name[2];
	int align_strand[2] = {0, 0};
	int seq_strand[2] = {0, 0};            // mSplitSeqStrand
	int seq_start[2] = {0, 0};             // mSplitAlignSeqStart (after FastaIndex::Get clamped it)
	int seq_length[2] = {0, 0};            // mSplitAlignSeqLength (after FastaIndex::Get set it to the fetched length)
	std::string remainder[2];              // mSplitRemainderSeq
};

// SplitAlignmentTask::CalculateBreakRegion (tools/SplitAlignment.cpp:637-655)
inline void BreakRegion(int min_read, int max_read, int max_fragment, int align_start, int align_end, int strand, int& break_start,
                 int& break_length)
{
	const int align_len = align_end - align_start + 1;
	const int push = std::min(max_read, (int)(0.5 * align_len));
	break_length = max_fragment - align_len - min_read + 2 * push;
	break_start = (strand == kPlus) ? align_end - push + 1 : align_start + push - 1;
}

// SplitAlignmentTask::Initialize (tools/SplitAlignment.cpp:31-175), the parts dosplitalign's output depends on:
// the two window sequences and the mate regions (genome + every overlapping transcript)
inline bool InitializeTask(ClusterTask& task, int id, const std::vector<Location>& pair, const FastaIndex& reference,
                    const ExonModel& exons, double frag_mean, double frag_sd, int min_read, int max_read)
{
	task.fusion_id = id;
	const int min_fragment = (int)(frag_mean - 3 * frag_sd);
	const int max_fragment = (int)(frag_mean + 3 * frag_sd);
	if (pair.size() != 2)
	{
		std::cerr << "Error: Incorrect input for SplitAlignment::Calculate()" << std::endl;
		return false;
	}
	for (int end = 0; end <= 1; end++)
	{
		const Location& loc = pair[end];
		const int ref_strand = (end == 0) ? loc.strand : 1 - loc.strand;
		int break_start, break_length;
		BreakRegion(min_read, max_read, max_fragment, loc.start, loc.end, loc.strand, break_start, break_length);
		int seq_start, seq_length;
		if (loc.strand == kPlus)
		{
			seq_start = break_start - max_read;
			seq_length = break_length + max_read;
		}
		else
		{
			seq_start = break_start - break_length + 1;
			seq_length = break_length + max_read;
		}
		reference.Get(loc.ref_name, ref_strand, seq_start, seq_length, task.window[end]);
		task.align_ref_name[end] = loc.ref_name;
		task.align_strand[end] = loc.strand;
		task.seq_strand[end] = ref_strand;
		task.seq_start[end] = seq_start;
		task.seq_length[end] = seq_length;
		// remainder of the aligned region outside the window, for the predicted sequence (SplitAlignment.cpp:81-103)
		task.remainder[end].clear();
		if (loc.strand == kPlus)
		{
			if (loc.start < seq_start)
			{
				int r_start = loc.start, r_length = (seq_start - 1) - loc.start + 1;
				reference.Get(loc.ref_name, ref_strand, r_start, r_length, task.remainder[end]);
			}
		}
		else if (loc.end > seq_start + seq_length - 1)
		{
			int r_start = seq_start + seq_length, r_length = loc.end - (seq_start + seq_length) + 1;
			reference.Get(loc.ref_name, ref_strand, r_start, r_length, task.remainder[end]);
		}

		std::string chromosome, gene, transcript;
		int genome_strand, genome_break_start;
		if (ParseTranscriptId(loc.ref_name, gene, transcript) && exons.IsTranscript(transcript))
		{
			exons.TranscriptToGenome(transcript, loc.strand, break_start, chromosome, genome_strand, genome_break_start);
		}
		else
		{
			chromosome = loc.ref_name;
			genome_strand = loc.strand;
			genome_break_start = break_start;
		}
		const int mate_min = min_fragment - break_length - max_read + 1;
		const int mate_max = max_fragment - min_read;
		Region mate;
		if (genome_strand == kPlus)
		{
			mate.start = genome_break_start - mate_max;
			mate.end = genome_break_start - mate_min;
		}
		else
		{
			mate.start = genome_break_start + mate_min;
			mate.end = genome_break_start + mate_max;
		}
		Location genome_region;
		genome_region.ref_name = chromosome;
		genome_region.strand = genome_strand;
		genome_region.start = mate.start;
		genome_region.end = mate.end;
		task.mate_regions[end].push_back(genome_region);

		std::vector<std::string> transcripts;
		exons.RegionTranscripts(chromosome, mate, transcripts);
		for (const std::string& tr : transcripts)
		{
			const std::string transcript_id = exons.GeneOf(tr) + "|" + tr;
			int r_start = 0, r_end = 0, r_strand = 0;
			if (exons.ThroughTranscript(tr, genome_break_start, 1 - genome_strand, mate_min, mate_max, r_strand, r_start, r_end))
			{
				Location l;
				l.ref_name = transcript_id;
				l.strand = 1 - r_strand;
				l.start = r_start;
				l.end = r_end;
				task.mate_regions[end].push_back(l);
			}
		}
	}
	return true;
}

// BinnedLocations (tools/SplitAlignment.cpp:177-229): 2000-bp bins per (strand, reference)
class BinnedRegions
{
public:
	explicit BinnedRegions(int spacing) : mSpacing(spacing) {}
	void Add(int id, const Location& loc)
	{
		const int idx = (int)mIds.size();
		mIds.push_back(id);
		mRegions.push_back(Region{loc.start, loc.end});
		const int first = loc.start / mSpacing, last = loc.end / mSpacing;
		for (int bin = first; bin <= last; bin++) mBinned[loc.strand][loc.ref_name][bin].push_back(idx);
	}
	void Overlapping(const std::string& ref, int strand, const Region& r, std::unordered_set<int>& ids) const
	{
		auto ref_it = mBinned[strand].find(ref);
		if (ref_it == mBinned[strand].end()) return;
		const int first = r.start / mSpacing, last = r.end / mSpacing;
		for (int bin = first; bin <= last; bin++)
		{
			auto b = ref_it->second.find(bin);
			if (b == ref_it->second.end()) continue;
			for (int idx : b->second)
			{
				const Region& g = mRegions[idx];
				if (g.start <= r.end && g.end >= r.start) ids.insert(mIds[idx]);
			}
		}
	}

private:
	int mSpacing;
	std::unordered_map<std::string, std::unordered_map<int, std::vector<int>>> mBinned[2];
	std::vector<int> mIds;
	std::vector<Region> mRegions;
};

struct PairHash
{
	// std::hash<pair<int,int>> of the oracle build = boost-style hash_combine; only lookups depend on it
	size_t operator()(const std::pair<int, int>& p) const
	{
		size_t seed = 0;
		seed ^= std::hash<int>()(p.first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
		seed ^= std::hash<int>()(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
		return seed;
	}
};
}  // namespace dfbhost

#endif
