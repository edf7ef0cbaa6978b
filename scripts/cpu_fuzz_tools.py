#!/usr/bin/env python
"""Differential fuzzing of the tools' input handling, without a GPU: small synthetic file sets are perturbed (odd but
legal records, and malformed ones) and handed to our tool (device double preloaded) and to the compiled reference tool;
exit code, output bytes and messages must agree.  Where the reference dies on a signal (an escaped bad_lexical_cast
aborts it) any non-zero exit of ours is accepted -- we report and exit 1 by design; where a DebugCheck of the reference
trips (exit 1, the message is its own source line) the exit code and the outputs must agree, not the wording.
Usage: python scripts/cpu_fuzz_tools.py <dosplitalign|matealign|localalign|evalsplitalign|splitseq> <seed> <seconds> [max rounds]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import files  # noqa: E402
import oracle  # noqa: E402  (the compiled reference tools)

BIN = os.path.join(ROOT, "defuse_b200", "bin")


def build_double(out):
    obj, lib = os.path.join(out, "dp_oracle.o"), os.path.join(out, "libdevice_double.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-c", os.path.join(ROOT, "oracle", "dp_oracle.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", lib,
                    os.path.join(ROOT, "tests", "device_double", "device_double.cpp"), obj, "-lpthread"], check=True)
    return lib


def mutate_lines(rng, text, ops, n_ops):
    """text -> text with n_ops random line-level edits drawn from ops; returns (text, [descriptions])."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    done = []
    for _ in range(n_ops):
        if not lines:
            break
        k = int(rng.integers(0, len(lines)))
        name, fn = ops[int(rng.integers(0, len(ops)))]
        try:
            out = fn(rng, lines[k])
        except (IndexError, ValueError):  # an edit that does not apply to a line an earlier edit already mangled
            continue
        done.append("%s@%d" % (name, k))
        if out is None:
            del lines[k]
        elif isinstance(out, list):
            lines[k:k + 1] = out
        else:
            lines[k] = out
    return "\n".join(lines) + "\n", done


def _field(idx, fn):
    def op(rng, line):
        f = line.split("\t")
        if len(f) <= idx:
            return line
        f[idx] = fn(rng, f[idx])
        return "\t".join(f)
    return op


SAM_OPS = [
    ("dup", lambda r, l: [l, l]),
    ("del", lambda r, l: None),
    ("qname_noslash", _field(0, lambda r, v: v.split("/")[0])),
    ("qname_flags", lambda r, l: "\t".join([l.split("\t")[0].split("/")[0], str(int(r.choice([0x40, 0x80, 0x50, 0x90, 0xC0])))] + l.split("\t")[2:])),
    ("qname_end3", _field(0, lambda r, v: v.split("/")[0] + "/3")),
    ("qname_twoslash", _field(0, lambda r, v: v + "/1")),
    ("qname_alpha", _field(0, lambda r, v: "x" + v)),
    ("flag_bits", _field(1, lambda r, v: str((int(v) if v.lstrip("-").isdigit() else 0) | int(r.choice([0x100, 0x400, 0x1, 0x20]))))),
    ("flag_bad", _field(1, lambda r, v: v + "x")),
    ("flag_neg", _field(1, lambda r, v: "-" + v)),
    ("rname_star", _field(2, lambda r, v: "*")),
    ("rname_unknown", _field(2, lambda r, v: v + "_nope")),
    ("pos_small", _field(3, lambda r, v: str(int(r.integers(-5, 3))))),
    ("pos_huge", _field(3, lambda r, v: str(int(r.integers(10 ** 6, 2 * 10 ** 9))))),
    ("pos_bad", _field(3, lambda r, v: v + ".5")),
    ("pos_plus", _field(3, lambda r, v: "+" + v)),
    ("seq_star", _field(9, lambda r, v: "*")),
    ("seq_long", _field(9, lambda r, v: v * int(r.integers(2, 6)))),
    ("seq_empty", _field(9, lambda r, v: "")),
    ("cut_fields", lambda r, l: "\t".join(l.split("\t")[:int(r.integers(1, 10))])),
    ("ten_fields", lambda r, l: "\t".join(l.split("\t")[:10])),
    ("extra_fields", lambda r, l: l + "\tXX:i:1\tYY:Z:abc"),
    ("trailing_tab", lambda r, l: l + "\t"),
    ("empty_line", lambda r, l: [l, ""]),
    ("header_mid", lambda r, l: ["@SQ\tSN:x\tLN:5", l]),
    ("crlf", lambda r, l: l + "\r"),
]

FASTQ_OPS = [  # applied to whole 4-line records: the mutator below hands over "l1\x00l2\x00l3\x00l4"
    ("del_read", lambda r, rec: None),
    ("dup_other_seq", lambda r, rec: [rec, rec.split("\0")[0] + "\0" + "ACGT" * 20 + "\0+\0" + "I" * 80]),
    ("lower", lambda r, rec: "\0".join([rec.split("\0")[0], rec.split("\0")[1].lower()] + rec.split("\0")[2:])),
    ("short", lambda r, rec: "\0".join([rec.split("\0")[0], rec.split("\0")[1][:int(r.integers(0, 30))], "+", "I"])),
    ("noslash", lambda r, rec: "\0".join([rec.split("\0")[0].split("/")[0]] + rec.split("\0")[1:])),
    ("end3", lambda r, rec: "\0".join([rec.split("\0")[0].split("/")[0] + "/3"] + rec.split("\0")[1:])),
    ("alpha_name", lambda r, rec: "\0".join(["@r" + rec.split("\0")[0][1:]] + rec.split("\0")[1:])),
    ("no_at", lambda r, rec: "\0".join([rec.split("\0")[0][1:]] + rec.split("\0")[1:])),
    ("drop_qual", lambda r, rec: "\0".join(rec.split("\0")[:3])),
    ("swap_end", lambda r, rec: "\0".join([rec.split("\0")[0][:-1] + ("2" if rec.split("\0")[0].endswith("1") else "1")] + rec.split("\0")[1:])),
]

REGION_OPS = [
    ("dup", lambda r, l: [l, l]),
    ("del", lambda r, l: None),
    ("second_region", lambda r, l: [l, "\t".join(l.split("\t")[:4] + [str(int(l.split("\t")[4]) + int(r.integers(-300, 300))), str(int(l.split("\t")[5]) + int(r.integers(-100, 600)))])]),
    ("flip_strand", _field(3, lambda r, v: "-" if v == "+" else "+")),
    ("start_small", _field(4, lambda r, v: str(int(r.integers(-50, 2))))),
    ("end_huge", _field(5, lambda r, v: str(int(v) + int(r.integers(1000, 100000))))),
    ("start_gt_end", lambda r, l: "\t".join(l.split("\t")[:4] + [l.split("\t")[5], l.split("\t")[4]])),
    ("big_id", _field(0, lambda r, v: str(int(v) + 100000))),
    ("bad_strand", _field(3, lambda r, v: "x")),
    ("bad_int", _field(4, lambda r, v: v + "a")),
    # (not 5 fields: the reference reads the sixth after checking for five, Parsers.cpp:237,251 -- undefined there)
    ("cut", lambda r, l: "\t".join(l.split("\t")[:int(r.integers(1, 5))])),
    ("unknown_chrom", _field(2, lambda r, v: v + "_nope")),
]


EXON_OPS = [
    ("dup", lambda r, l: [l, l]),
    ("del", lambda r, l: None),
    ("odd_coords", lambda r, l: l + "\t" + str(int(r.integers(1, 5000)))),
    ("more_exons", lambda r, l: l + "\t%d\t%d" % (int(l.split("\t")[-1]) + 50, int(l.split("\t")[-1]) + 120)),
    ("unknown_chrom", _field(2, lambda r, v: v + "_nope")),
    ("flip_strand", _field(3, lambda r, v: "-" if v == "+" else "+")),
    ("bad_strand", _field(3, lambda r, v: "?")),
    ("bad_int", _field(4, lambda r, v: v + "q")),
    ("cut", lambda r, l: "\t".join(l.split("\t")[:int(r.integers(1, 6))])),
    ("same_transcript", _field(1, lambda r, v: "T0_0")),
    ("empty_line", lambda r, l: [l, ""]),
    ("shift", lambda r, l: "\t".join(l.split("\t")[:4] + [str(int(v) + 7) for v in l.split("\t")[4:]])),
]


def mutate_fastq(rng, text, n_ops):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    recs = ["\0".join(lines[k:k + 4]) for k in range(0, len(lines), 4)]
    joined, done = mutate_lines(rng, "\n".join(recs) + "\n", FASTQ_OPS, n_ops)
    out = joined.replace("\0", "\n")
    if rng.random() < 0.15:
        out = out[:-1]  # no newline at the end of the file
    return out, done


def run(cmd, stdin=None, env=None):
    p = subprocess.run(cmd, input=stdin, capture_output=True, timeout=300, env=env)
    err = b"".join(l for l in p.stderr.splitlines(keepends=True) if not l.startswith(b"[fai_load]"))
    return p.returncode, p.stdout, err


def fuzz_dosplitalign(rng, d, env, rnd):
    sub = os.path.join(d, "s%d" % rnd)
    kw = dict(seed=int(rng.integers(1, 10 ** 6)), n_clusters=int(rng.integers(2, 9)), pairs_per_cluster=int(rng.integers(2, 12)),
              L=int(rng.choice([60, 76, 100])), frag_mean=int(rng.choice([180, 250])))
    if rng.random() < 0.3:
        kw.update(read_len_jitter=int(rng.integers(1, 9)), lower_frac=0.02)
    args = files.make_split_dataset(sub, **kw)
    what = []
    for name, ops, p_mut in (("improper.sam", SAM_OPS, 0.8), ("clusters.regions", REGION_OPS, 0.4), ("exons.regions", EXON_OPS, 0.25)):
        if rng.random() < p_mut:
            path = os.path.join(sub, name)
            text, done = mutate_lines(rng, open(path).read(), ops, int(rng.integers(1, 4)))
            open(path, "w").write(text)
            what += [name + ":" + x for x in done]
    for name in ("reads.1.fastq", "reads.2.fastq"):
        if rng.random() < 0.35:
            path = os.path.join(sub, name)
            text, done = mutate_fastq(rng, open(path).read(), int(rng.integers(1, 4)))
            open(path, "w").write(text)
            what += [name + ":" + x for x in done]
    ours_out, ref_out = os.path.join(sub, "ours.tmp"), os.path.join(sub, "ref.tmp")
    ro = run([os.path.join(BIN, "dosplitalign")] + args + ["-a", ours_out], env=env)
    rr = run([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", ref_out])
    fo = open(ours_out, "rb").read() if os.path.exists(ours_out) else None
    fr = open(ref_out, "rb").read() if os.path.exists(ref_out) else None
    verdict = compare(ro, rr, fo, fr)
    if verdict:
        keep = os.path.join(ROOT, "gpurun_out", "fuzz_fail_dosplitalign_%d" % rnd)
        shutil.rmtree(keep, ignore_errors=True)
        shutil.copytree(sub, keep)
    shutil.rmtree(sub, ignore_errors=True)
    return verdict, kw, what, ro, rr


LOCAL_OPS = [
    ("dup", lambda r, l: [l, l]),
    ("del", lambda r, l: None),
    ("cut", lambda r, l: "\t".join(l.split("\t")[:int(r.integers(1, 3))])),
    ("extra", lambda r, l: l + "\tjunk\tmore"),
    ("empty_seq", _field(2, lambda r, v: "")),
    ("empty_ref", _field(1, lambda r, v: "")),
    ("lower_seq", _field(2, lambda r, v: v.lower())),
    ("n_seq", _field(2, lambda r, v: v[:len(v) // 2] + "NNNN" + v[len(v) // 2:])),
    ("odd_bytes", _field(2, lambda r, v: v[:3] + "-*x." + v[3:])),
    ("short_ref", _field(1, lambda r, v: v[:int(r.integers(0, 40))])),
    ("empty_line", lambda r, l: [l, ""]),
    ("crlf", lambda r, l: l + "\r"),
    ("spaces", lambda r, l: l.replace("\t", " ", 1)),
    ("empty_id", _field(0, lambda r, v: "")),
]


def fuzz_localalign(rng, d, env, rnd):
    kw = dict(seed=int(rng.integers(1, 10 ** 6)), n_refs=int(rng.integers(1, 5)), n_lines=int(rng.integers(1, 40)),
              R=int(rng.choice([170, 301, 2001])), L=(int(rng.integers(1, 40)), int(rng.integers(40, 160))))
    text = files.make_localalign_input(**kw).decode()
    what = []
    if rng.random() < 0.8:
        text, what = mutate_lines(rng, text, LOCAL_OPS, int(rng.integers(1, 4)))
    if rng.random() < 0.1:
        text = text[:-1]
    m = int(rng.integers(1, 12))
    args = ["-m", str(m), "-x", str(int(rng.integers(-8, 1))), "-g", str(int(rng.integers(-8, 1)))]
    if rng.random() < 0.7:
        args += ["-t", "%.2f" % rng.random()]
    ro = run([os.path.join(BIN, "localalign")] + args, text.encode(), env)
    rr = run([oracle.ref_tool("ref_localalign")] + args, text.encode())
    verdict = compare(ro, rr, None, None)
    if verdict:
        open(os.path.join(ROOT, "gpurun_out", "fuzz_fail_localalign_%d.txt" % rnd), "w").write(" ".join(args) + "\n" + text)
    return verdict, dict(kw, args=args), what, ro, rr


def fuzz_matealign(rng, d, env, rnd):
    sub = os.path.join(d, "m%d" % rnd)
    kw = dict(seed=int(rng.integers(1, 10 ** 6)), n_pairs=int(rng.integers(2, 40)), L=int(rng.choice([36, 76, 150])),
              search=int(rng.choice([200, 400, 1000])))
    args, sam = files.make_matealign_dataset(sub, **kw)
    sam = sam.decode()
    what = []
    if rng.random() < 0.8:
        sam, done = mutate_lines(rng, sam, SAM_OPS, int(rng.integers(1, 4)))
        what += ["sam:" + x for x in done]
    for name in ("reads.1.fastq", "reads.2.fastq"):
        if rng.random() < 0.35:
            path = os.path.join(sub, name)
            text, done = mutate_fastq(rng, open(path).read(), int(rng.integers(1, 4)))
            open(path, "w").write(text)
            what += [name + ":" + x for x in done]
    ro = run([os.path.join(BIN, "matealign")] + args, sam.encode(), env)
    rr = run([oracle.ref_tool("ref_matealign")] + args, sam.encode())
    verdict = compare(ro, rr, None, None)
    if verdict:
        keep = os.path.join(ROOT, "gpurun_out", "fuzz_fail_matealign_%d" % rnd)
        shutil.rmtree(keep, ignore_errors=True)
        shutil.copytree(sub, keep)
        open(os.path.join(keep, "in.sam"), "w").write(sam)
        open(os.path.join(keep, "args.txt"), "w").write("\n".join(args))
    shutil.rmtree(sub, ignore_errors=True)
    return verdict, kw, what, ro, rr


ALIGN_OPS = [
    ("dup", lambda r, l: [l, l]),
    ("del", lambda r, l: None),
    ("score", _field(8, lambda r, v: str(int(r.integers(-40, 200))))),
    ("score_bad", _field(8, lambda r, v: v + "x")),
    ("ref_split_shift", _field(4, lambda r, v: str(int(v) + int(r.integers(-3, 4))))),
    ("ref_split2_shift", _field(5, lambda r, v: str(int(v) + int(r.integers(-3, 4))))),
    ("ref_split_far", _field(4, lambda r, v: str(int(r.integers(-5, 2000))))),
    ("ref_split2_far", _field(5, lambda r, v: str(int(r.integers(-5, 2000))))),
    ("read_split", lambda r, l: "\t".join(l.split("\t")[:6] + [str(int(r.integers(0, 120))), str(int(r.integers(0, 120)))] + l.split("\t")[8:])),
    ("read_split_anchor", lambda r, l: "\t".join(l.split("\t")[:6] + ["4", "4"] + l.split("\t")[8:])),
    ("revcomp_bad", _field(3, lambda r, v: "2")),
    ("fragment_bad", _field(1, lambda r, v: "f" + v)),
    ("fusion_other", _field(0, lambda r, v: str(int(v) + int(r.integers(1, 3))))),
    ("fusion_unknown", _field(0, lambda r, v: str(int(v) + 5000))),
    ("fusion_bad", _field(0, lambda r, v: v + "z")),
    ("cut", lambda r, l: "\t".join(l.split("\t")[:int(r.integers(1, 9))])),
    ("no_trailing_tab", lambda r, l: l.rstrip("\t")),
    ("extra", lambda r, l: l + "more\tfields"),
    ("empty_line", lambda r, l: [l, ""]),
]


def fuzz_evalsplitalign(rng, d, env, rnd):
    sub = os.path.join(d, "e%d" % rnd)
    kw = dict(seed=int(rng.integers(1, 10 ** 6)), n_clusters=int(rng.integers(2, 9)), pairs_per_cluster=int(rng.integers(4, 16)),
              L=int(rng.choice([60, 76, 100])))
    args = files.make_split_dataset(sub, **kw)
    raw, srt = os.path.join(sub, "raw.alignments"), os.path.join(sub, "sorted.alignments")
    subprocess.run([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", raw], check=True, capture_output=True)
    files.sort_alignments(raw, srt)
    what = []
    if rng.random() < 0.85:
        text, what = mutate_lines(rng, open(srt).read(), ALIGN_OPS, int(rng.integers(1, 5)))
        open(srt, "w").write(text if rng.random() < 0.9 else text[:-1])
    common, ev = files.downstream_args(args, sub)
    outs = lambda tag: ["-q", os.path.join(sub, tag + ".seq"), "-b", os.path.join(sub, tag + ".break"), "-p", os.path.join(sub, tag + ".pred")]
    tenv = dict(env)
    if rng.random() < 0.5:  # regions of a few lines on several threads
        tenv.update(DFB_TOOL_CHUNK_MIN=str(int(rng.integers(1, 400))), DFB_TOOL_THREADS=str(int(rng.integers(1, 9))))
    ro = run([os.path.join(BIN, "evalsplitalign")] + ev + outs("ours"), env=tenv)
    rr = run([oracle.ref_tool("ref_evalsplitalign")] + ev + outs("ref"))
    # the prediction records are buffered by the reference and lost when it exits early: compared on success only
    names = ["seq", "break", "pred"] if rr[0] == 0 else ["seq", "break"]
    if b"Unable to find max score split" in rr[2]:
        names.remove("break")  # the reference writes uninitialised break positions for such a fusion (SplitAlignment.cpp:530-536)
    import re
    # (the empty prediction of a fusion without regions is labelled with the id of a default-constructed task:
    # uninitialised in the reference -- 0 in most runs, anything in others)
    mask = lambda b: re.sub(rb"(?m)^-?\d+(\tN\t0\t0\t-1\t-1)$", rb"?\1", b)
    fo = b"|".join(mask(open(os.path.join(sub, "ours." + k), "rb").read()) for k in names)
    fr = b"|".join(mask(open(os.path.join(sub, "ref." + k), "rb").read()) for k in names)
    verdict = compare(ro, rr, fo, fr)
    if not verdict and rr[0] > 0 and fo != fr:
        verdict = "outputs in front of the error differ"
    if verdict:
        keep = os.path.join(ROOT, "gpurun_out", "fuzz_fail_evalsplitalign_%d" % rnd)
        shutil.rmtree(keep, ignore_errors=True)
        shutil.copytree(sub, keep)
        open(os.path.join(keep, "args.txt"), "w").write("\n".join(ev))
    shutil.rmtree(sub, ignore_errors=True)
    return verdict, kw, what, ro, rr


def fuzz_splitseq(rng, d, env, rnd):
    sub = os.path.join(d, "q%d" % rnd)
    kw = dict(seed=int(rng.integers(1, 10 ** 6)), n_clusters=int(rng.integers(2, 7)), pairs_per_cluster=int(rng.integers(4, 14)),
              L=int(rng.choice([60, 76, 100])))
    if rng.random() < 0.3:
        kw.update(read_len_jitter=int(rng.integers(1, 9)), lower_frac=0.02, n_rate=0.01)
    args = files.make_split_dataset(sub, **kw)
    raw, srt = os.path.join(sub, "raw.alignments"), os.path.join(sub, "sorted.alignments")
    subprocess.run([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", raw], check=True, capture_output=True)
    files.sort_alignments(raw, srt)
    common, ev = files.downstream_args(args, sub)
    pred = os.path.join(sub, "ref.pred")
    subprocess.run([oracle.ref_tool("ref_evalsplitalign")] + ev + ["-q", os.path.join(sub, "ref.seq"), "-b", os.path.join(sub, "ref.break"), "-p", pred],
                   check=True, capture_output=True)
    prefix = files.write_read_index(sub)
    source = pred if rng.random() < 0.5 else srt  # the predicted records, or every record
    what = []
    if rng.random() < 0.7:
        ops = [o for o in ALIGN_OPS if o[0] in ("dup", "del", "ref_split_shift", "ref_split2_shift", "score", "read_split", "revcomp_bad",
                                                 "fragment_bad", "fusion_other", "fusion_unknown", "cut", "no_trailing_tab", "empty_line")]
        text, what = mutate_lines(rng, open(source).read(), ops, int(rng.integers(1, 4)))
        open(source, "w").write(text)
    tail = common + ["-p", prefix, "-a", source]
    if rng.random() < 0.3:
        tail += ["-i", str(int(rng.integers(0, kw["n_clusters"] + 2)))]
    ro = run([os.path.join(BIN, "splitseq")] + tail, env=env)
    rr = run([oracle.ref_tool("ref_splitseq")] + tail)
    verdict = compare(ro, rr, None, None)
    if verdict:
        keep = os.path.join(ROOT, "gpurun_out", "fuzz_fail_splitseq_%d" % rnd)
        shutil.rmtree(keep, ignore_errors=True)
        shutil.copytree(sub, keep)
        open(os.path.join(keep, "args.txt"), "w").write("\n".join(tail))
    shutil.rmtree(sub, ignore_errors=True)
    return verdict, kw, what, ro, rr


def compare(ro, rr, fo, fr):
    """'' when the two runs agree, else what differs."""
    if rr[0] < 0:  # the reference died on a signal
        return "" if ro[0] != 0 else "reference died on signal %d, ours exited 0" % -rr[0]
    if ro[0] != rr[0]:
        return "exit codes differ: ours %d, reference %d" % (ro[0], rr[0])
    if b" failed on line: " in rr[2]:  # a DebugCheck of the reference (message = its source line): ours words it differently
        return "" if ro[1] == rr[1] else "stdout differs"
    if rr[0] == 0 and fo != fr:
        return "output files differ"
    if ro[1] != rr[1]:
        return "stdout differs"
    if ro[2] != rr[2]:
        return "stderr differs"
    return ""


def main():
    tool, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    max_rounds = int(sys.argv[4]) if len(sys.argv) > 4 else None
    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    stats = {"tool": tool, "seed": seed, "rounds": 0, "reference_exit_0": 0, "reference_exit_1": 0, "reference_signal": 0, "disagreements": []}
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, LD_PRELOAD=build_double(d))
        while time.time() < t_end and (max_rounds is None or stats["rounds"] < max_rounds):
            verdict, kw, what, ro, rr = {"dosplitalign": fuzz_dosplitalign, "matealign": fuzz_matealign, "localalign": fuzz_localalign, "evalsplitalign": fuzz_evalsplitalign, "splitseq": fuzz_splitseq}[tool](rng, d, env, stats["rounds"])
            stats["rounds"] += 1
            stats["reference_exit_0" if rr[0] == 0 else ("reference_signal" if rr[0] < 0 else "reference_exit_1")] += 1
            if verdict:
                stats["disagreements"].append({"round": stats["rounds"] - 1, "what": verdict, "dataset": kw, "mutations": what,
                                               "ours": [ro[0], ro[2].decode(errors="replace")[-300:]],
                                               "reference": [rr[0], rr[2].decode(errors="replace")[-300:]]})
                if len(stats["disagreements"]) >= 10:
                    break
    print(json.dumps(stats, indent=1))
    return 1 if stats["disagreements"] else 0


if __name__ == "__main__":
    sys.exit(main())
