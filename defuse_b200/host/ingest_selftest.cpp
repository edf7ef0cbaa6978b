// ingest_selftest -- the parallel, in-place FASTQ index (fast_io.h) against the line-by-line FastqReader
// (host_common.h) that mirrors tools/ReadStream.cpp: same reads, same last-one-wins map, same message where the
// stream ends early.  Usage: ingest_selftest <file.fastq> <threads>   (prints one line per check, exit 1 on a mismatch)
#include "host_common.h"
#include "fast_io.h"

#include <map>

using namespace dfbhost;

// --blocks <bytes>: stdin through LineBlocks back to stdout; exit 1 when a block that is not the last one does not
// end behind a newline
static int BlocksMode(size_t block_bytes)
{
	LineBlocks in;
	in.Open(0, block_bytes);
	const char* p = nullptr;
	size_t n = 0;
	bool open_tail = false; // a block without a final newline must be the last
	size_t blocks = 0;
	while (in.Next(p, n))
	{
		if (open_tail) return 1;
		if (n == 0) return 1;
		open_tail = p[n - 1] != '\n';
		fwrite(p, 1, n, stdout);
		blocks++;
	}
	fprintf(stderr, "blocks %zu\n", blocks);
	return 0;
}

int main(int argc, char* argv[])
{
	if (argc >= 3 && std::string(argv[1]) == "--blocks") return BlocksMode((size_t)atoll(argv[2]));
	if (argc < 3) return 2;
	const std::string path = argv[1];
	const int threads = atoi(argv[2]);
	// sequential view
	std::map<int, std::string> want;
	std::string want_message;
	bool want_fatal = false;
	{
		std::ostringstream captured;
		std::streambuf* old = std::cerr.rdbuf(captured.rdbuf());
		FastqReader reader;
		if (!reader.Open(path)) return 2;
		FastqRead rd;
		while (reader.Next(rd))
		{
			int fragment = 0;
			if (!ParseInt(rd.fragment, fragment))
			{
				captured << "Error: bad lexical cast: fragment index '" << rd.fragment << "'" << std::endl;
				want_fatal = true;
				break;
			}
			want[PackId(fragment, rd.read_end)] = rd.sequence;
		}
		std::cerr.rdbuf(old);
		want_message = captured.str();
	}
	FastqIndex index;
	if (!index.Open(path)) return 2;
	index.Scan(threads);
	int bad = 0;
	if (index.Message() != want_message || index.Fatal() != want_fatal)
	{
		std::cout << "message mismatch: got [" << index.Message() << "] fatal " << index.Fatal() << " want [" << want_message
		          << "] fatal " << want_fatal << std::endl;
		bad++;
	}
	size_t found = 0;
	for (const auto& kv : want)
	{
		const char* seq = nullptr;
		uint32_t len = 0;
		if (!index.Find(kv.first, seq, len) || std::string(seq, len) != kv.second)
		{
			if (bad < 5) std::cout << "read " << kv.first << " differs" << std::endl;
			bad++;
		}
		else
			found++;
	}
	// nothing beyond the sequential view: probe ids around the known ones
	for (const auto& kv : want)
	{
		const char* seq = nullptr;
		uint32_t len = 0;
		const int other = kv.first ^ (int)0x80000000;
		if (want.find(other) == want.end() && index.Find(other, seq, len))
		{
			if (bad < 5) std::cout << "read " << other << " should be absent" << std::endl;
			bad++;
		}
	}
	std::cout << "reads " << want.size() << " matched " << found << " message [" << want_message.substr(0, 60) << "] bad " << bad << std::endl;
	return bad ? 1 : 0;
}
