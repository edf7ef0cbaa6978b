// fast_io.h -- host ingest of the drop-in tools: whole inputs mapped into memory, cut into
// line-aligned chunks that worker threads parse without allocating, then merged in file order so that every tool
// still sees exactly the sequence of records (and the first error) a line-by-line reader would.
#ifndef DFB_FAST_IO_H
#define DFB_FAST_IO_H

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace dfbhost
{
// threads a tool uses for parsing / formatting (DFB_TOOL_THREADS overrides)
inline int ToolThreads()
{
	static int n = 0;
	if (n == 0)
	{
		const char* e = getenv("DFB_TOOL_THREADS");
		n = e ? atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
		if (n < 1) n = 1;
	}
	return n;
}

// fn(0..T-1) on T threads, the caller being thread 0
template <class F>
inline void ParallelRun(int T, F fn)
{
	if (T <= 0) return;
	if (T == 1)
	{
		fn(0);
		return;
	}
	std::vector<std::thread> th;
	th.reserve((size_t)T - 1);
	for (int k = 1; k < T; k++) th.emplace_back(fn, k);
	fn(0);
	for (auto& t : th) t.join();
}

// A whole input in memory: regular files are mapped, anything else (stdin, pipes) is read to the end.
class MappedInput
{
public:
	MappedInput() = default;
	MappedInput(const MappedInput&) = delete;
	MappedInput& operator=(const MappedInput&) = delete;
	~MappedInput() { Close(); }

	bool OpenFile(const std::string& path)
	{
		const int fd = open(path.c_str(), O_RDONLY);
		if (fd < 0) return false;
		const bool ok = FromFd(fd);
		close(fd);
		return ok;
	}
	bool OpenStdin() { return FromFd(0); }
	const char* data() const { return mData; }
	size_t size() const { return mSize; }

private:
	bool FromFd(int fd)
	{
		Close();
		struct stat st;
		if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode))
		{
			mSize = (size_t)st.st_size;
			if (mSize == 0) return true;
			void* p = mmap(nullptr, mSize, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
			if (p != MAP_FAILED)
			{
				mData = (const char*)p;
				mMapped = true;
				madvise(p, mSize, MADV_SEQUENTIAL);
				return true;
			}
			mSize = 0;
		}
		// stream: read to the end
		size_t cap = 1 << 24;
		char* buf = (char*)malloc(cap);
		if (!buf) return false;
		size_t n = 0;
		for (;;)
		{
			if (n == cap)
			{
				cap *= 2;
				char* nb = (char*)realloc(buf, cap);
				if (!nb)
				{
					free(buf);
					return false;
				}
				buf = nb;
			}
			const ssize_t got = read(fd, buf + n, cap - n);
			if (got < 0)
			{
				if (errno == EINTR) continue;
				free(buf);
				return false;
			}
			if (got == 0) break;
			n += (size_t)got;
		}
		mData = buf;
		mSize = n;
		mMapped = false;
		return true;
	}
	void Close()
	{
		if (mData)
		{
			if (mMapped) munmap((void*)mData, mSize); else free((void*)mData);
		}
		mData = nullptr;
		mSize = 0;
	}
	const char* mData = nullptr;
	size_t mSize = 0;
	bool mMapped = false;
};

// [begin, end) holds whole lines; first_line = number of lines in front of it (0-based)
struct LineChunk
{
	size_t begin, end;
	int64_t first_line;
};

// Cuts [0, n) into at most T chunks that start right behind a '\n' and counts the lines in front of each
// (a last line without '\n' counts when it is not empty -- std::getline's view of the file).
inline std::vector<LineChunk> SplitLines(const char* p, size_t n, int T)
{
	std::vector<LineChunk> chunks;
	if (n == 0) return chunks;
	static const size_t min_chunk = getenv("DFB_TOOL_CHUNK_MIN") ? (size_t)std::max(1, atoi(getenv("DFB_TOOL_CHUNK_MIN"))) : (size_t)1 << 16;
	T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, n / min_chunk + 1)); // (tests shrink the chunks to a few lines)
	std::vector<size_t> cut((size_t)T + 1, n);
	cut[0] = 0;
	for (int k = 1; k < T; k++)
	{
		size_t at = n / (size_t)T * (size_t)k;
		if (at < cut[(size_t)k - 1]) at = cut[(size_t)k - 1];
		const char* nl = at < n ? (const char*)memchr(p + at, '\n', n - at) : nullptr;
		cut[(size_t)k] = nl ? (size_t)(nl - p) + 1 : n;
	}
	std::vector<int64_t> lines((size_t)T, 0);
	ParallelRun(T, [&](int k) {
		const char* a = p + cut[(size_t)k];
		const char* const e = p + cut[(size_t)k + 1];
		int64_t c = 0;
		while (a < e)
		{
			const char* nl = (const char*)memchr(a, '\n', (size_t)(e - a));
			if (!nl) break; // (only the last chunk can end without a newline; that tail is one more line for its parser)
			c++;
			a = nl + 1;
		}
		lines[(size_t)k] = c;
	});
	int64_t first = 0;
	for (int k = 0; k < T; k++)
	{
		if (cut[(size_t)k] < cut[(size_t)k + 1]) chunks.push_back(LineChunk{cut[(size_t)k], cut[(size_t)k + 1], first});
		first += lines[(size_t)k];
	}
	return chunks;
}

// lexical_cast<int> on [b, e): optional sign, digits only, must fit an int
inline bool ParseIntRange(const char* b, const char* e, int& out)
{
	if (b >= e) return false;
	bool neg = false;
	if (*b == '-' || *b == '+')
	{
		neg = *b == '-';
		b++;
	}
	if (b >= e) return false;
	long long v = 0;
	for (; b < e; b++)
	{
		if (*b < '0' || *b > '9') return false;
		v = v * 10 + (*b - '0');
		if (v > 4294967296LL) return false;
	}
	if (neg) v = -v;
	if (v < -2147483648LL || v > 2147483647LL) return false;
	out = (int)v;
	return true;
}

inline void AppendInt(std::string& s, long long v)
{
	char buf[24];
	auto r = std::to_chars(buf, buf + sizeof(buf), v);
	s.append(buf, (size_t)(r.ptr - buf));
}

// Set of 64-bit keys, open addressing; only membership matters to its users.
class KeySet
{
public:
	explicit KeySet(size_t expected = 1024)
	{
		size_t cap = 1024;
		while (cap < expected * 2) cap <<= 1;
		mSlots.assign(cap, kEmpty);
	}
	// true when the key was not present yet
	bool Insert(uint64_t key)
	{
		if (key == kEmpty)
		{
			const bool fresh = !mHasAllOnes;
			mHasAllOnes = true;
			return fresh;
		}
		if ((mCount + 1) * 10 > mSlots.size() * 7) Grow();
		return Put(mSlots, key) ? (mCount++, true) : false;
	}

private:
	static constexpr uint64_t kEmpty = ~0ull; // marks a free slot; the key of that value is tracked by a flag
	static bool Put(std::vector<uint64_t>& slots, uint64_t key)
	{
		const size_t mask = slots.size() - 1;
		size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & mask;
		for (;;)
		{
			if (slots[h] == kEmpty)
			{
				slots[h] = key;
				return true;
			}
			if (slots[h] == key) return false;
			h = (h + 1) & mask;
		}
	}
	void Grow()
	{
		std::vector<uint64_t> bigger(mSlots.size() * 2, kEmpty);
		for (uint64_t k : mSlots)
			if (k != kEmpty) Put(bigger, k);
		mSlots.swap(bigger);
	}
	std::vector<uint64_t> mSlots;
	size_t mCount = 0;
	bool mHasAllOnes = false;
};

// ---------------------------------------------------------------------------------------------
// FASTQ: "@<fragment>/<1|2>", four lines per read (tools/ReadStream.cpp:59-103), indexed in place
// ---------------------------------------------------------------------------------------------
// What a sequential FastqReadStream + SplitReadRealigner::AddReads (tools/SplitAlignment.cpp:253-264) would leave in
// mReads, as views into the mapped file: read id -> sequence line.  A malformed name ends the stream there with the
// reference's message (reads in front of it are kept); a fragment name that is not an integer is fatal.
class FastqIndex
{
public:
	// returns false when the file cannot be used at all (extension / open), like IReadStream::Create
	bool Open(const std::string& filename)
	{
		const std::string::size_type dot = filename.find_last_of('.');
		const std::string ext = filename.substr(dot + 1);
		if (ext != "fastq" && ext != "fq")
		{
			mMessage = "Error: unrecognized extension " + ext + "\n";
			return false;
		}
		if (!mInput.OpenFile(filename))
		{
			mMessage = "Error: unable to open file " + filename + "\n";
			return false;
		}
		return true;
	}

	// Parses the whole file on T threads.  Afterwards: Message() = what to print on stderr ("" if nothing),
	// Fatal() = whether the reference would have died at that point.
	// keep_order: also keep the reads as a list in file order (Count / Record), for tools that walk the file
	void Scan(int T, bool keep_order = false)
	{
		const char* p = mInput.data();
		const size_t n = mInput.size();
		std::vector<LineChunk> chunks = SplitLines(p, n, T);
		struct Part
		{
			std::vector<int> id;
			std::vector<uint64_t> off;
			std::vector<uint32_t> len;
			int64_t event_record = -1; // first record that ends the stream
			bool fatal = false;
			std::string message;
		};
		std::vector<Part> parts(chunks.size());
		ParallelRun((int)chunks.size(), [&](int k) {
			const LineChunk& c = chunks[(size_t)k];
			Part& part = parts[(size_t)k];
			const char* a = p + c.begin;
			const char* const e = p + c.end;
			// records are 4 lines from the top of the file: skip the tail of a record that began in an earlier chunk
			int64_t line = c.first_line;
			const char* ls[4];
			const char* le[4];
			int have = 0;
			auto next_line = [&](const char*& b, const char*& en) -> bool {
				if (a >= e) return false;
				const char* nl = (const char*)memchr(a, '\n', (size_t)(e - a));
				b = a;
				en = nl ? nl : e;
				a = nl ? nl + 1 : e;
				line++;
				return true;
			};
			const char *b, *en;
			while (line % 4 != 0)
				if (!next_line(b, en)) return;
			part.id.reserve((size_t)(e - a) / 200 + 16);
			part.off.reserve(part.id.capacity());
			part.len.reserve(part.id.capacity());
			for (;;)
			{
				const int64_t record = line / 4;
				have = 0;
				const char* save = a;
				while (have < 4 && a < e && next_line(ls[have], le[have])) have++;
				if (have < 4)
				{
					// the record continues in the next chunk (or the file ends inside it)
					if (k + 1 < (int)chunks.size())
					{
						const char* q = p + c.end;
						const char* const fe = p + n;
						while (have < 4 && q < fe)
						{
							const char* nl = (const char*)memchr(q, '\n', (size_t)(fe - q));
							ls[have] = q;
							le[have] = nl ? nl : fe;
							q = nl ? nl + 1 : fe;
							have++;
						}
					}
					if (have < 4) return; // fewer than four lines left: GetNextRead returns false
					(void)save;
				}
				const char* h = ls[0];
				const char* he = le[0];
				if (h >= he || *h != '@')
				{
					part.event_record = record;
					part.message = "Error: Unable to interpret read name " + std::string(h, he) + "\n";
					return;
				}
				const char* slash = (const char*)memchr(h, '/', (size_t)(he - h));
				const char end_name = (!slash || slash + 1 >= he) ? '\0' : slash[1];
				if (end_name != '1' && end_name != '2')
				{
					part.event_record = record;
					part.message = "Error: Unable to interpret read end " + std::string(h, he) + "\n";
					return;
				}
				int fragment = 0;
				if (!ParseIntRange(h + 1, slash, fragment))
				{
					part.event_record = record;
					part.fatal = true;
					part.message = "Error: bad lexical cast: fragment index '" + std::string(h + 1, slash) + "'\n";
					return;
				}
				part.id.push_back((int)(((unsigned)fragment & 0x7fffffffu) | ((unsigned)(end_name == '1' ? 0 : 1) << 31)));
				part.off.push_back((uint64_t)(ls[1] - p));
				part.len.push_back((uint32_t)(le[1] - ls[1]));
				if (a >= e) return;
			}
		});
		// merge in file order up to the first event
		size_t total = 0;
		size_t last_part = parts.size();
		for (size_t k = 0; k < parts.size(); k++)
		{
			total += parts[k].id.size();
			if (parts[k].event_record >= 0)
			{
				mMessage = parts[k].message;
				mFatal = parts[k].fatal;
				last_part = k + 1;
				break;
			}
		}
		unsigned max_index = 0;
		for (size_t k = 0; k < last_part && k < parts.size(); k++)
			for (int id : parts[k].id) max_index = std::max(max_index, (unsigned)id & 0x7fffffffu);
		if (keep_order)
		{
			for (size_t k = 0; k < last_part && k < parts.size(); k++)
			{
				mOrderId.insert(mOrderId.end(), parts[k].id.begin(), parts[k].id.end());
				mOrderOff.insert(mOrderOff.end(), parts[k].off.begin(), parts[k].off.end());
				mOrderLen.insert(mOrderLen.end(), parts[k].len.begin(), parts[k].len.end());
			}
			return;
		}
		mDense = (size_t)max_index < 8 * total + 4096;
		if (mDense)
		{
			mOff.assign(2 * ((size_t)max_index + 1), kAbsent);
			mLen.assign(2 * ((size_t)max_index + 1), 0);
		}
		for (size_t k = 0; k < last_part && k < parts.size(); k++)
		{
			const Part& part = parts[k];
			for (size_t q = 0; q < part.id.size(); q++)
			{
				if (mDense)
				{
					const size_t slot = 2 * (size_t)((unsigned)part.id[q] & 0x7fffffffu) + ((unsigned)part.id[q] >> 31);
					mOff[slot] = part.off[q];
					mLen[slot] = part.len[q];
				}
				else
				{
					mSparse[part.id[q]] = std::make_pair(part.off[q], part.len[q]);
				}
			}
		}
	}

	const std::string& Message() const { return mMessage; }
	bool Fatal() const { return mFatal; }

	// file-order view (Scan(T, true)): read k is PackId(fragment, end) = id, sequence [seq, seq + len)
	size_t Count() const { return mOrderId.size(); }
	void Record(size_t k, int& id, const char*& seq, uint32_t& len) const
	{
		id = mOrderId[k];
		seq = mInput.data() + mOrderOff[k];
		len = mOrderLen[k];
	}

	// the sequence of read `id` (PackId(fragment, end)), or false when this file does not hold it
	bool Find(int id, const char*& seq, uint32_t& len) const
	{
		if (mDense)
		{
			const size_t slot = 2 * (size_t)((unsigned)id & 0x7fffffffu) + ((unsigned)id >> 31);
			if (slot >= mOff.size() || mOff[slot] == kAbsent) return false;
			seq = mInput.data() + mOff[slot];
			len = mLen[slot];
			return true;
		}
		auto it = mSparse.find(id);
		if (it == mSparse.end()) return false;
		seq = mInput.data() + it->second.first;
		len = it->second.second;
		return true;
	}

private:
	static constexpr uint64_t kAbsent = ~0ull;
	MappedInput mInput;
	bool mDense = true;
	std::vector<uint64_t> mOff;
	std::vector<uint32_t> mLen;
	std::unordered_map<int, std::pair<uint64_t, uint32_t>> mSparse;
	std::vector<int> mOrderId;
	std::vector<uint64_t> mOrderOff;
	std::vector<uint32_t> mOrderLen;
	std::string mMessage;
	bool mFatal = false;
};

// Whole-line blocks of an input that may be too large to hold at once (localalign reads gigabytes from stdin): a
// mapped file is handed out in windows, a pipe is read block by block on a helper thread while the previous block
// is being processed.  Every block ends behind a '\n' except possibly the last.
class LineBlocks
{
public:
	LineBlocks() = default;
	LineBlocks(const LineBlocks&) = delete;
	LineBlocks& operator=(const LineBlocks&) = delete;
	~LineBlocks()
	{
		if (mPrefetch.joinable()) mPrefetch.join();
		if (mMap) munmap((void*)mMap, mMapSize);
		free(mBuf[0].p);
		free(mBuf[1].p);
	}
	void Open(int fd, size_t block_bytes)
	{
		mFd = fd;
		mBlock = std::max<size_t>(block_bytes, 16);
		struct stat st;
		if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0)
		{
			const off_t at = lseek(fd, 0, SEEK_CUR);
			void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (p != MAP_FAILED)
			{
				mMap = (const char*)p;
				mMapSize = (size_t)st.st_size;
				mMapPos = at > 0 ? (size_t)at : 0;
				madvise(p, mMapSize, MADV_SEQUENTIAL);
				return;
			}
		}
		StartRead(0, 0);
	}
	// next block of whole lines; false at the end of the input.  The memory stays valid until the next call.
	bool Next(const char*& p, size_t& n)
	{
		if (mMap)
		{
			if (mMapPos >= mMapSize) return false;
			size_t end = std::min(mMapSize, mMapPos + mBlock);
			if (end < mMapSize)
			{
				const char* nl = (const char*)memchr(mMap + end - 1, '\n', mMapSize - end + 1);
				end = nl ? (size_t)(nl - mMap) + 1 : mMapSize;
			}
			p = mMap + mMapPos;
			n = end - mMapPos;
			mMapPos = end;
			return true;
		}
		for (;;)
		{
			if (mPrefetch.joinable()) mPrefetch.join();
			if (mFinished) return false;
			Buf& cur = mBuf[mFill];
			const size_t have = mHave;
			if (mEof)
			{
				mFinished = true;
				p = cur.p;
				n = have;
				return have > 0;
			}
			size_t cut = have;
			while (cut > 0 && cur.p[cut - 1] != '\n') cut--;
			if (cut == 0)
			{
				// one line longer than the block: keep reading into the same buffer
				StartRead(mFill, have);
				continue;
			}
			const int other = 1 - mFill;
			mBuf[other].Ensure(have - cut + mBlock);
			if (have > cut) memcpy(mBuf[other].p, cur.p + cut, have - cut);
			p = cur.p;
			n = cut;
			StartRead(other, have - cut);
			return true;
		}
	}

private:
	struct Buf
	{
		char* p = nullptr;
		size_t cap = 0;
		void Ensure(size_t n)
		{
			if (n <= cap) return;
			char* q = (char*)realloc(p, n);
			if (!q)
			{
				fprintf(stderr, "Error: out of memory reading the input\n");
				exit(1);
			}
			p = q;
			cap = n;
		}
	};
	// reads up to one more block into mBuf[which] behind `carry` bytes, on a helper thread
	void StartRead(int which, size_t carry)
	{
		mFill = which;
		mBuf[which].Ensure(carry + mBlock);
		mPrefetch = std::thread([this, carry, which] {
			size_t have = carry;
			const size_t want = carry + mBlock;
			char* const buf = mBuf[which].p;
			while (have < want)
			{
				const ssize_t got = read(mFd, buf + have, want - have);
				if (got < 0 && errno == EINTR) continue;
				if (got <= 0)
				{
					mEof = true;
					break;
				}
				have += (size_t)got;
			}
			mHave = have;
		});
	}
	int mFd = 0;
	size_t mBlock = 1 << 28;
	const char* mMap = nullptr;
	size_t mMapSize = 0, mMapPos = 0;
	Buf mBuf[2];
	int mFill = 0;
	size_t mHave = 0;
	bool mEof = false, mFinished = false;
	std::thread mPrefetch;
};

}  // namespace dfbhost

#endif
