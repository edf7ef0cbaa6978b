#!/usr/bin/env python
"""Generates tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on seeded inputs.  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no golden vectors for this path, so these files are the
pinned outputs of the reference itself: per-task results of SplitReadAligner / SimpleAligner through
oracle/ref_harness.cpp, and whole-tool outputs of ref_localalign."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import util  # noqa: E402


def s(b):
    return b.decode("latin-1")


def split_cases():
    rng = np.random.default_rng(101)
    cases = []
    # planted junctions, dosplitalign scoring
    refs, reads, tc, tr = util.split_batch(rng, 6, 5, (30, 110), 80, 360, sub=0.02, indel=0.005, n_rate=0.01)
    for c, r in zip(tc, tr):
        cases.append(dict(read=reads[r], ref1=refs[2 * c], ref2=refs[2 * c + 1], params=[2, -1, -2, 0, 8], thr="ms"))
    # adversarial: ties, poly-A, N, lowercase, empty, windows shorter than the read
    adv = [
        (b"A" * 40, b"A" * 60, b"A" * 50), (b"ACGT" * 8, b"ACGT" * 20, b"TTGCA" * 12), (b"", b"ACGTACGT", b"GGGTTT"),
        (b"ACGTNNNNACGTacgtNNNNACGT", b"ACGTTGCANNNNACGT" * 4, b"acgtNNNNACGTTTGA" * 4),
        (b"GATTACA" * 5, b"GATTACAGATTACA", b"CATCATCAT"), (b"N" * 20, b"ACGTN" * 10, b"NNNNNACGT" * 5),
        (b"ACGTACGTACGTAAAAAAAAAAAAAAAAAAAA", b"TTTTACGTACGTACGTCCCC", b"GGGAAAAAAAAAAAAAAAAAAAATTT"),
        (b"T", b"T", b"T"), (b"ACGTACGTAC", b"", b"ACGTACGTAC"), (b"ACGTACGTAC", b"ACGTACGTAC", b""),
    ]
    for read, r1, r2 in adv:
        for thr in ("ms", 0):
            cases.append(dict(read=read, ref1=r1, ref2=r2, params=[2, -1, -2, 0, 8], thr=thr))
    # other constructor arguments (endGaps, minSplitScore <= 0, positive gap ...)
    rng = np.random.default_rng(102)
    refs, reads, tc, tr = util.split_batch(rng, 3, 3, (6, 30), 10, 50, sub=0.05)
    for params in ([2, -1, -2, 1, 8], [2, -1, -2, 0, 0], [2, -1, -2, 0, -3], [1, -1, 1, 0, 4], [3, 1, -2, 1, 0],
                   [300, -200, -250, 0, 1200], [0, 0, 0, 0, 0]):
        for c, r in zip(tc, tr):
            cases.append(dict(read=reads[r], ref1=refs[2 * c], ref2=refs[2 * c + 1], params=params, thr="ms"))
            cases.append(dict(read=reads[r], ref1=refs[2 * c], ref2=refs[2 * c + 1], params=params, thr=0))
    out = []
    for c in cases:
        m, x, g, eg, ms = c["params"]
        thr = int(np.float64(np.float32(np.float32(len(c["read"])) * np.float32(m))) * 0.90) if c["thr"] == "ms" else int(c["thr"])
        n, best = oracle.split_align_count(c["read"], c["ref1"], c["ref2"], thr, m, x, g, eg, ms, impl="ref")
        entry = dict(read=s(c["read"]), ref1=s(c["ref1"]), ref2=s(c["ref2"]), params=c["params"], min_score=thr,
                     n_alignments=n, best=best)
        if n <= 400:
            al = oracle.split_align(c["read"], c["ref1"], c["ref2"], thr, m, x, g, bool(eg), ms, impl="ref")
            entry["alignments"] = al.tolist()
        out.append(entry)
    return out


def simple_cases():
    rng = np.random.default_rng(103)
    out = []
    refs, seqs, tr, ts = util.simple_batch(rng, 6, 40, (1, 400), (1, 150))
    pairs = [(refs[a], seqs[b]) for a, b in zip(tr, ts)]
    pairs += [(b"", b""), (b"A", b""), (b"", b"A"), (b"ACGTN", b"GTN"), (b"ACGT", b"acgt"), (b"A" * 50, b"A" * 20),
              (bytes(range(1, 120)), bytes(range(40, 90))), (b"GATTACA" * 10, b"TACAGATT")]
    for scoring in ([10, -5, -5], [2, -1, -2], [1, 0, 0], [1, -1, 1], [0, 0, 0], [-1, -2, -3], [300, -200, -250], [3, 3, 3]):
        for ref, seq in pairs:
            out.append(dict(ref=s(ref), seq=s(seq), scoring=scoring,
                            score=oracle.simple_align(ref, seq, *scoring, impl="ref")))
    return out


def localalign_case():
    """Whole-tool golden: stdin lines -> ref_localalign stdout (tools/localalign.cpp)."""
    rng = np.random.default_rng(104)
    refs, seqs, tr, ts = util.simple_batch(rng, 5, 60, 300, (40, 80), related=0.7)
    lines = ["id%d\t%s\t%s" % (k, s(refs[a]), s(seqs[b])) for k, (a, b) in enumerate(zip(tr, ts))]
    lines.append("empty\tACGT\t")          # empty sequence: percent is -nan
    lines.append("extra\tACGTACGT\tCGTA\tignored-field")
    text = "\n".join(lines) + "\n"
    tool = oracle.ref_tool("ref_localalign")
    res = {}
    for name, args in (("t0.8", ["-m", "10", "-x", "-5", "-g", "-5", "-t", "0.8"]), ("nothreshold", ["-m", "2", "-x", "-1", "-g", "-2"])):
        p = subprocess.run([tool] + args, input=text.encode(), stdout=subprocess.PIPE, check=True)
        res[name] = dict(args=args, stdout=p.stdout.decode())
    return dict(stdin=text, runs=res)


def tool_cases():
    """Whole-tool goldens on synthetic files (synth/files.py): stdout / -a file of the compiled reference tools."""
    import tempfile
    from synth import files
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, kw in (("split_small", dict(seed=1, n_clusters=30, pairs_per_cluster=25)),
                         ("split_jitter_lower", dict(seed=7, n_clusters=20, pairs_per_cluster=25, read_len_jitter=20,
                                                     lower_frac=0.01, n_rate=0.01))):
            sub = os.path.join(d, name)
            args = files.make_split_dataset(sub, **kw)
            res = os.path.join(sub, "ref.alignments")
            subprocess.run([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", res], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[name] = dict(kw=kw, output=open(res).read())
        sub = os.path.join(d, "mate")
        args, sam = files.make_matealign_dataset(sub, seed=4, n_pairs=120)
        p = subprocess.run([oracle.ref_tool("ref_matealign")] + args, input=sam, stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, check=True)
        out["mate_small"] = dict(kw=dict(seed=4, n_pairs=120), output=p.stdout.decode())
    return out


def downstream_cases():
    """evalsplitalign and splitseq of the compiled reference on the sorted output of ref_dosplitalign."""
    import tempfile
    from synth import files
    kw = dict(seed=5, n_clusters=24, pairs_per_cluster=40)
    with tempfile.TemporaryDirectory() as d:
        args = files.make_split_dataset(d, **kw)
        raw = os.path.join(d, "ref.alignments")
        subprocess.run([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", raw], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        files.sort_alignments(raw, os.path.join(d, "sorted.alignments"))
        common, ev = files.downstream_args(args, d)
        names = {k: os.path.join(d, "ref." + k) for k in ("seq", "break", "predalign")}
        subprocess.run([oracle.ref_tool("ref_evalsplitalign")] + ev + ["-q", names["seq"], "-b", names["break"],
                                                                       "-p", names["predalign"]], check=True)
        out = dict(kw=kw, sorted=open(os.path.join(d, "sorted.alignments")).read())
        for k, v in names.items():
            out[k] = open(v).read()
        prefix = files.write_read_index(d)
        base = common + ["-p", prefix, "-a", names["predalign"]]
        out["splitseq"] = subprocess.run([oracle.ref_tool("ref_splitseq")] + base, stdout=subprocess.PIPE, check=True).stdout.decode()
        ids = sorted({int(l.split("\t")[0]) for l in out["predalign"].splitlines()})
        out["splitseq_id"] = {str(i): subprocess.run([oracle.ref_tool("ref_splitseq")] + base + ["-i", str(i)],
                                                     stdout=subprocess.PIPE, check=True).stdout.decode() for i in ids[:2]}
    return out


if __name__ == "__main__":
    assert oracle.have_ref(), "build the reference first: make -C oracle ref"
    json.dump(split_cases(), open(os.path.join(HERE, "split_aligner.json"), "w"), indent=0)
    json.dump(simple_cases(), open(os.path.join(HERE, "simple_aligner.json"), "w"), indent=0)
    json.dump(localalign_case(), open(os.path.join(HERE, "localalign_tool.json"), "w"), indent=0)
    json.dump(tool_cases(), open(os.path.join(HERE, "tools.json"), "w"), indent=0)
    json.dump(downstream_cases(), open(os.path.join(HERE, "tools_downstream.json"), "w"), indent=0)
    print("golden vectors written to", HERE)
