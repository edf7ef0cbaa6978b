"""Host ingest (defuse_b200/host/fast_io.h): the parallel in-place FASTQ index must equal a line-by-line reader that
mirrors tools/ReadStream.cpp -- for well-formed files and for every way a stream can end early -- whatever the number
of threads and wherever the chunk boundaries fall.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "defuse_b200", "bin", "ingest_selftest")


def _records(rng, n, start=0):
    out = []
    for k in range(n):
        L = int(rng.integers(0, 130))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
        qual = b"@" * L if k % 7 == 0 else (b"+" * L if k % 11 == 0 else b"I" * L)   # quality lines that look like headers
        out.append(b"@%d/%d\n%s\n+\n%s\n" % (start + k, 1 + k % 2, seq, qual))
    return out


CASES = {
    "plain": lambda rng: b"".join(_records(rng, 400)),
    "no_trailing_newline": lambda rng: b"".join(_records(rng, 57))[:-1],
    "truncated_record": lambda rng: b"".join(_records(rng, 80)) + b"@999/1\nACGT\n",
    "duplicates_last_wins": lambda rng: b"".join(_records(rng, 50) + _records(rng, 50)),
    "bad_name_in_the_middle": lambda rng: b"".join(_records(rng, 120)) + b"X12/1\nAC\n+\nII\n" + b"".join(_records(rng, 90, 500)),
    "bad_end_in_the_middle": lambda rng: b"".join(_records(rng, 33)) + b"@77/3\nAC\n+\nII\n" + b"".join(_records(rng, 90, 500)),
    "no_slash": lambda rng: b"".join(_records(rng, 20)) + b"@77\nAC\n+\nII\n" + b"".join(_records(rng, 10, 500)),
    "fragment_not_an_int": lambda rng: b"".join(_records(rng, 64)) + b"@read7/1\nAC\n+\nII\n" + b"".join(_records(rng, 30, 500)),
    "sparse_ids": lambda rng: b"".join(b"@%d/1\nACGT\n+\nIIII\n" % (k * 1000003 % 2000000011) for k in range(1, 200)),
    "empty": lambda rng: b"",
    "blank_lines_shift_records": lambda rng: b"".join(_records(rng, 10)) + b"\n" + b"".join(_records(rng, 10, 100)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_fastq_index_equals_sequential_reader(tmp_path, name):
    if not os.path.exists(TOOL):
        pytest.skip("tools not built")
    rng = np.random.default_rng(sum(map(ord, name)))
    path = str(tmp_path / "r.fastq")
    open(path, "wb").write(CASES[name](rng))
    for threads, chunk in ((1, "65536"), (3, "64"), (8, "700"), (16, "1")):
        p = subprocess.run([TOOL, path, str(threads)], capture_output=True, env=dict(os.environ, DFB_TOOL_CHUNK_MIN=chunk))
        assert p.returncode == 0, (name, threads, chunk, p.stdout.decode()[-600:], p.stderr.decode()[-300:])


def test_dosplitalign_sam_errors_are_the_first_in_file_order(tmp_path):
    """Chunks of SAM lines are parsed in parallel; the tool must still die on the line a sequential reader would
    have died on, with the reference's message (tools/AlignmentStream.cpp:51-62,100-105).  No GPU needed: the
    errors come before any alignment."""
    import sys
    sys.path.insert(0, ROOT)
    from synth import files
    tool = os.path.join(ROOT, "defuse_b200", "bin", "dosplitalign")
    if not os.path.exists(tool):
        pytest.skip("tools not built")
    d = str(tmp_path / "d")
    args = files.make_split_dataset(d, seed=3, n_clusters=6, pairs_per_cluster=30)
    good = open(os.path.join(d, "improper.sam")).read().splitlines()
    assert len(good) > 100
    cases = [
        (good[:40] + [""] + good[40:90] + ["too\tfew\tfields"] + good[90:], "Error: Empty alignment line 41"),
        (good[:70] + ["too\tfew\tfields"] + good[70:80] + [""] + good[80:], "Error: Format error for alignment line 71"),
        (good[:25] + ["7/3\t0\tchr1\t100\t255\t50M\t*\t0\t0\tACGT\tIIII"] + good[25:] + [""],
         "Error: Unable to interpret qname for alignment line 26"),
    ]
    from oracle import ref_tool
    ref = ref_tool("ref_dosplitalign")
    for k, (lines, message) in enumerate(cases):
        sam = os.path.join(d, "bad%d.sam" % k)
        open(sam, "w").write("\n".join(lines) + "\n")
        a = [x if x != os.path.join(d, "improper.sam") else sam for x in args] + ["-a", os.path.join(d, "out.tmp")]
        for chunk in ("1", "300", "65536"):
            p = subprocess.run([tool] + a, capture_output=True, env=dict(os.environ, DFB_TOOL_CHUNK_MIN=chunk, DFB_TOOL_THREADS="8"))
            assert p.returncode == 1 and p.stderr.decode().strip().splitlines()[-1] == message, (k, chunk, p.stderr.decode()[-300:])
        if ref:
            q = subprocess.run([ref] + a, capture_output=True)
            assert q.returncode == 1 and message in q.stderr.decode(), q.stderr.decode()[-300:]


@pytest.mark.parametrize("mode", ["pipe", "file"])
def test_line_blocks_reproduce_the_input(tmp_path, mode):
    """LineBlocks (what localalign reads stdin through): whole-line blocks, read ahead from a pipe or windows of a
    mapped file; concatenated they are the input, whatever the block size -- including lines longer than a block and
    an input without a final newline."""
    if not os.path.exists(TOOL):
        pytest.skip("tools not built")
    rng = np.random.default_rng(5)
    lines = [bytes(rng.integers(65, 91, int(n)).astype(np.uint8)) for n in rng.integers(0, 400, 300)]
    lines[17] = b"x" * 9000                      # longer than most block sizes below
    for text in (b"\n".join(lines) + b"\n", b"\n".join(lines), b"", b"\n", b"no newline at all"):
        path = str(tmp_path / "in.txt")
        open(path, "wb").write(text)
        for block in (1, 16, 333, 4096, 1 << 20):
            if mode == "pipe":
                p = subprocess.run([TOOL, "--blocks", str(block)], input=text, capture_output=True)
            else:
                with open(path, "rb") as f:
                    p = subprocess.run([TOOL, "--blocks", str(block)], stdin=f, capture_output=True)
            assert p.returncode == 0 and p.stdout == text, (mode, block, len(text), p.stderr[-200:])
