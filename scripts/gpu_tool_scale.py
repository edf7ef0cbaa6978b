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


def localalign_scale(n_lines, n_refs):
    """BASELINE.json configs[1]: localalign, 100-bp reads against 2001-bp references, input redirected from a file."""
    import numpy as np
    out = {"lines": n_lines, "refs": n_refs}
    sc = ["-m", "10", "-x", "-5", "-g", "-5", "-t", "0.8"]
    with tempfile.TemporaryDirectory() as d:
        rng = np.random.default_rng(2)
        refs = [files.rand_seq(rng, 2001) for _ in range(n_refs)]
        path = os.path.join(d, "in.txt")
        t0 = time.perf_counter()
        sample_lines = max(2000, n_lines // 100)
        with open(path, "wb") as f, open(path + ".sample", "wb") as fs:
            per_ref = max(1, n_lines // n_refs)
            k = 0
            for r in refs:                         # the pipeline repeats a reference on consecutive lines
                offs = rng.integers(0, 2001 - 100, per_ref)
                unrelated = rng.random(per_ref) < 0.2
                buf = []
                for o, u in zip(offs, unrelated):
                    seq = files.rand_seq(rng, 100) if u else files.mutate(rng, r[o:o + 100], 0.02, 0.0)
                    buf.append(b"c%d\t%s\t%s\n" % (k, r, seq))
                    k += 1
                blob = b"".join(buf)
                f.write(blob)
                if k <= sample_lines:
                    fs.write(blob)
        out["generate_s"] = time.perf_counter() - t0
        out["stdin_MB"] = os.path.getsize(path) / 1e6
        runs = []
        for rep in range(3):
            t0 = time.perf_counter()
            with open(path, "rb") as f:
                p = subprocess.run([os.path.join(BIN, "localalign")] + sc, stdin=f, capture_output=True, env=dict(os.environ, DFB_TRACE="1"))
            runs.append(time.perf_counter() - t0)
            assert p.returncode == 0, p.stderr.decode()[-2000:]
        out["ours_s"], out["ours_runs_s"] = min(runs), runs
        out["lines_per_s"] = k / out["ours_s"]
        out["gcups_wall"] = k * 2001 * 100 / out["ours_s"] / 1e9
        out["trace_last_run"] = [l for l in p.stderr.decode().splitlines() if l.startswith("[tool]")]
        ours_full = p.stdout
        # through a pipe as the pipeline does it (cat | localalign)
        t0 = time.perf_counter()
        p2 = subprocess.run("cat %s | %s %s" % (path, os.path.join(BIN, "localalign"), " ".join(sc)), shell=True, capture_output=True)
        out["ours_pipe_s"] = time.perf_counter() - t0
        out["pipe_identical"] = p2.stdout == ours_full
        ref = oracle.ref_tool("ref_localalign")
        if ref:
            n_sample = sum(1 for _ in open(path + ".sample", "rb"))
            t0 = time.perf_counter()
            with open(path + ".sample", "rb") as f:
                q = subprocess.run([ref] + sc, stdin=f, capture_output=True)
            ref_s = time.perf_counter() - t0
            with open(path + ".sample", "rb") as f:
                o = subprocess.run([os.path.join(BIN, "localalign")] + sc, stdin=f, capture_output=True)
            out["reference_sample"] = {"lines": n_sample, "seconds": ref_s, "lines_per_s": n_sample / ref_s,
                                       "identical": q.stdout == o.stdout}
            out["speedup_in_lines_per_s"] = out["lines_per_s"] / out["reference_sample"]["lines_per_s"]
    return out


def matealign_scale(n_pairs):
    """BASELINE.json configs[3]: matealign, 150-bp pairs, search length 1000."""
    out = {"pairs": n_pairs}
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        margs, sam = files.make_matealign_dataset(os.path.join(d, "m"), seed=4, n_pairs=n_pairs)
        out["generate_s"] = time.perf_counter() - t0
        runs = []
        for rep in range(3):
            t0 = time.perf_counter()
            p = subprocess.run([os.path.join(BIN, "matealign")] + margs, input=sam, capture_output=True, env=dict(os.environ, DFB_TRACE="1"))
            runs.append(time.perf_counter() - t0)
            assert p.returncode == 0, p.stderr.decode()[-2000:]
        out["ours_s"], out["ours_runs_s"] = min(runs), runs
        out["records"] = p.stdout.count(b"\n")
        out["pairs_per_s"] = n_pairs / out["ours_s"]
        out["trace_last_run"] = [l for l in p.stderr.decode().splitlines() if l.startswith("[tool]")]
        ref = oracle.ref_tool("ref_matealign")
        if ref:
            n_sample = max(2000, n_pairs // 50)
            sargs, ssam = files.make_matealign_dataset(os.path.join(d, "s"), seed=4, n_pairs=n_sample)
            t0 = time.perf_counter()
            q = subprocess.run([ref] + sargs, input=ssam, capture_output=True)
            ref_s = time.perf_counter() - t0
            o = subprocess.run([os.path.join(BIN, "matealign")] + sargs, input=ssam, capture_output=True)
            out["reference_sample"] = {"pairs": n_sample, "seconds": ref_s, "pairs_per_s": n_sample / ref_s, "identical": q.stdout == o.stdout}
            out["speedup_in_pairs_per_s"] = out["pairs_per_s"] / out["reference_sample"]["pairs_per_s"]
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "localalign":
        print(json.dumps(localalign_scale(int(sys.argv[2]) if len(sys.argv) > 2 else 1000000, int(sys.argv[3]) if len(sys.argv) > 3 else 10000)))
    elif len(sys.argv) > 1 and sys.argv[1] == "matealign":
        print(json.dumps(matealign_scale(int(sys.argv[2]) if len(sys.argv) > 2 else 300000)))
    else:
        main()
