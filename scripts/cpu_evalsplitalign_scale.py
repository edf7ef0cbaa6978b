#!/usr/bin/env python
"""evalsplitalign at scale (host-only tool, no GPU needed): a dosplitalign-shaped dataset whose sorted records are
replicated with fresh fragment indices to millions of lines; our tool against the compiled reference tool on the same
file, outputs compared byte for byte.
Usage: python scripts/cpu_evalsplitalign_scale.py [n_clusters] [replication] > profiles/<tag>_tool_scale_evalsplitalign.json"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import files  # noqa: E402
import oracle  # noqa: E402  (baseline side only)

BIN = os.path.join(ROOT, "defuse_b200", "bin")


def timed(cmd, reps):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True, capture_output=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def main():
    n_clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    ref_split, ref_eval = oracle.ref_tool("ref_dosplitalign"), oracle.ref_tool("ref_evalsplitalign")
    with tempfile.TemporaryDirectory() as d:
        args = files.make_split_dataset(d, seed=71, n_clusters=n_clusters, pairs_per_cluster=8, n_chrom=8, genes_per_chrom=40)
        subprocess.run([ref_split] + args + ["-a", os.path.join(d, "raw.alignments")], check=True, capture_output=True)
        files.sort_alignments(os.path.join(d, "raw.alignments"), os.path.join(d, "small.alignments"))
        n_lines = 0
        with open(os.path.join(d, "sorted.alignments"), "w") as out:
            run, cur = [], None

            def flush():
                n = 0
                for k in range(rep):
                    for f in run:
                        out.write("\t".join([f[0], str(int(f[1]) + 100000 * k)] + f[2:]) + "\t\n")
                        n += 1
                return n
            for line in open(os.path.join(d, "small.alignments")):
                f = line.rstrip("\n").split("\t")[:9]
                if f[0] != cur:
                    n_lines += flush()
                    run, cur = [], f[0]
                run.append(f)
            n_lines += flush()
        common, ev = files.downstream_args(args, d)
        outs = lambda tag: ["-q", os.path.join(d, tag + ".seq"), "-b", os.path.join(d, tag + ".break"), "-p", os.path.join(d, tag + ".pred")]
        ours_s = timed([os.path.join(BIN, "evalsplitalign")] + ev + outs("ours"), 4)
        ours1_s = timed(["env", "DFB_TOOL_THREADS=1", os.path.join(BIN, "evalsplitalign")] + ev + outs("ours"), 2)
        ref_s = timed([ref_eval] + ev + outs("ref"), 1)
        same = all(open(os.path.join(d, "ours." + k), "rb").read() == open(os.path.join(d, "ref." + k), "rb").read()
                   for k in ("seq", "break", "pred"))
        print(json.dumps({"tool": "evalsplitalign", "host_cpus": os.cpu_count(), "fusions": n_clusters, "records": n_lines,
                          "input_bytes": os.path.getsize(os.path.join(d, "sorted.alignments")),
                          "ours_s": ours_s, "ours_one_thread_s": ours1_s, "reference_s": ref_s,
                          "records_per_s": n_lines / ours_s, "speedup": ref_s / ours_s, "identical": same,
                          "where": "container CPU (no GPU involved: the tool is host-only)"}))


if __name__ == "__main__":
    main()
