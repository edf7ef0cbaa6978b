#!/usr/bin/env python
"""Tool-level wall clock on the GPU box: the drop-in tools vs the compiled reference tools on the same synthetic
files (BASELINE.json configs[0]-shaped dosplitalign, localalign, matealign).  Checks byte identity, prints one JSON
object.  Usage (under gpurun): python scripts/gpu_tools_bench.py > gpurun_out/tools_bench.json"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import files  # noqa: E402
import oracle  # noqa: E402  (bench-side checker/baseline only)

BIN = os.path.join(ROOT, "defuse_b200", "bin")


def timed(cmd, stdin=None):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    dt = time.perf_counter() - t0
    assert p.returncode == 0, (cmd[0], p.stderr.decode()[-500:])
    return dt, p.stdout


def main():
    out = {"host_cpus": os.cpu_count()}
    with tempfile.TemporaryDirectory() as d:
        # dosplitalign: 200 clusters, 20 000 read pairs
        args = files.make_split_dataset(os.path.join(d, "s"), seed=1, n_clusters=200, pairs_per_cluster=100)
        ours, theirs = os.path.join(d, "s", "ours.tmp"), os.path.join(d, "s", "ref.tmp")
        timed([os.path.join(BIN, "dosplitalign")] + args + ["-a", ours])  # warm-up (module load, .fai)
        t_ours, _ = timed([os.path.join(BIN, "dosplitalign")] + args + ["-a", ours])
        t_ref, _ = timed([oracle.ref_tool("ref_dosplitalign")] + args + ["-a", theirs])
        a, b = open(ours).read(), open(theirs).read()
        out["dosplitalign"] = {"read_pairs": 20000, "clusters": 200, "records": len(b.splitlines()), "identical": a == b,
                               "ours_s": t_ours, "reference_s": t_ref, "speedup": t_ref / t_ours}
        # localalign: 40 000 lines, 2001-bp references
        text = files.make_localalign_input(seed=2, n_refs=400, n_lines=40000)
        sc = ["-m", "10", "-x", "-5", "-g", "-5", "-t", "0.8"]
        timed([os.path.join(BIN, "localalign")] + sc, text)
        t_ours, o1 = timed([os.path.join(BIN, "localalign")] + sc, text)
        t_ref, o2 = timed([oracle.ref_tool("ref_localalign")] + sc, text)
        out["localalign"] = {"lines": 40000, "identical": o1 == o2, "ours_s": t_ours, "reference_s": t_ref,
                             "speedup": t_ref / t_ours, "stdin_MB": len(text) / 1e6}
        # matealign: 20 000 pairs, 150 bp, search length 1000
        margs, sam = files.make_matealign_dataset(os.path.join(d, "m"), seed=4, n_pairs=20000)
        timed([os.path.join(BIN, "matealign")] + margs, sam)
        t_ours, o1 = timed([os.path.join(BIN, "matealign")] + margs, sam)
        t_ref, o2 = timed([oracle.ref_tool("ref_matealign")] + margs, sam)
        out["matealign"] = {"pairs": 20000, "identical": o1 == o2, "ours_s": t_ours, "reference_s": t_ref,
                            "speedup": t_ref / t_ours}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
