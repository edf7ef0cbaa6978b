#!/usr/bin/env python
"""Tool-level wall clock at scale (one fastq split of BASELINE.json configs[2]: ~1 M read pairs, 10 000 clusters) with
DFB_TRACE phase timings; the reference tool is timed on a 2 % subsample of the same generator and extrapolated in
read pairs.  Usage (under gpurun): python scripts/gpu_tool_scale.py [n_clusters] [pairs_per_cluster] > gpurun_out/tool_scale.json"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import files  # noqa: E402
import oracle  # noqa: E402  (bench-side baseline only)

BIN = os.path.join(ROOT, "defuse_b200", "bin")


def main():
    n_clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    ppc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    out = {"host_cpus": os.cpu_count(), "read_pairs": n_clusters * ppc, "clusters": n_clusters}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        args = files.make_split_dataset(os.path.join(d, "s"), seed=3, n_clusters=n_clusters, pairs_per_cluster=ppc,
                                        n_chrom=8, genes_per_chrom=40)
        out["generate_s"] = time.perf_counter() - t0
        res = os.path.join(d, "s", "ours.tmp")
        runs = []
        for rep in range(3):
            t0 = time.perf_counter()
            p = subprocess.run([os.path.join(BIN, "dosplitalign")] + args + ["-a", res], env=dict(os.environ, DFB_TRACE="1"),
                               capture_output=True)
            dt = time.perf_counter() - t0
            assert p.returncode == 0, p.stderr.decode()[-2000:]
            runs.append(dt)
            trace = p.stderr.decode()
        out["ours_s"] = min(runs)
        out["ours_runs_s"] = runs
        out["records"] = sum(1 for _ in open(res))
        out["read_pairs_per_s"] = out["read_pairs"] / out["ours_s"]
        out["trace_last_run"] = [l for l in trace.splitlines() if l.startswith("[tool]")]
        sys.stderr.write(trace)
        # the reference on a 2 % subsample of the same generator (first fastq split would take about an hour)
        sub_c = max(20, n_clusters // 50)
        sargs = files.make_split_dataset(os.path.join(d, "r"), seed=3, n_clusters=sub_c, pairs_per_cluster=ppc,
                                         n_chrom=8, genes_per_chrom=40)
        ref = oracle.ref_tool("ref_dosplitalign")
        if ref:
            rres, ores = os.path.join(d, "r", "ref.tmp"), os.path.join(d, "r", "ours.tmp")
            t0 = time.perf_counter()
            subprocess.run([ref] + sargs + ["-a", rres], check=True, capture_output=True)
            ref_s = time.perf_counter() - t0
            subprocess.run([os.path.join(BIN, "dosplitalign")] + sargs + ["-a", ores], check=True, capture_output=True)
            out["reference_sample"] = {"read_pairs": sub_c * ppc, "seconds": ref_s, "read_pairs_per_s": sub_c * ppc / ref_s,
                                       "identical": open(rres).read() == open(ores).read()}
            out["speedup_in_read_pairs_per_s"] = out["read_pairs_per_s"] / out["reference_sample"]["read_pairs_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
