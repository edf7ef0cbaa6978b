#!/usr/bin/env python
"""Host phases of the tools at scale, without a GPU: the real tool binaries with tests/device_double preloaded in
skip-DP mode (every DP answered "no alignment" at once), DFB_TRACE phase timers, minimum over five runs per phase.
What is timed is everything the host does around the GPU calls (ingest, candidates, tables, formatting of an empty
result) -- not the DP and not CUDA start-up.  With `--digest-against <binary dir>` the work handed to the device
library (task order, windows, reads, thresholds) is hashed for both builds and compared.
Usage: python scripts/cpu_tool_host_phases.py [--digest-against <dir with older tool binaries>] > profiles/<tag>_tool_host_phases.json"""
import collections
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import files  # noqa: E402

BIN = os.path.join(ROOT, "defuse_b200", "bin")


def build_double(out):
    obj, lib = os.path.join(out, "dp_oracle.o"), os.path.join(out, "libdevice_double.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-c", os.path.join(ROOT, "oracle", "dp_oracle.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", lib,
                    os.path.join(ROOT, "tests", "device_double", "device_double.cpp"), obj, "-lpthread"], check=True)
    return lib


def phases(cmd, stdin_path, env, runs=5):
    best = collections.OrderedDict()
    for _ in range(runs):
        with open(stdin_path, "rb") if stdin_path else open(os.devnull, "rb") as f:
            p = subprocess.run(cmd, stdin=f, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        assert p.returncode == 0, p.stderr.decode()[-1000:]
        for line in p.stderr.decode().splitlines():
            if line.startswith("[tool]"):
                name, ms = line[7:36].strip(), float(line[36:].split()[0])
                best[name] = min(best.get(name, ms), ms)
    return best


def digest(cmd, stdin_path, env, path):
    if os.path.exists(path):
        os.unlink(path)
    with open(stdin_path, "rb") if stdin_path else open(os.devnull, "rb") as f:
        subprocess.run(cmd, stdin=f, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=dict(env, DEVICE_DOUBLE_DIGEST=path), check=True)
    return open(path).read()


def main():
    other = sys.argv[sys.argv.index("--digest-against") + 1] if "--digest-against" in sys.argv else None
    out = {"host_cpus": os.cpu_count(), "where": "build container, device double in skip-DP mode (no GPU, no DP, no CUDA start-up)"}
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, LD_PRELOAD=build_double(d), DEVICE_DOUBLE_SKIP_DP="1", DFB_TRACE="1")
        sargs = files.make_split_dataset(os.path.join(d, "s"), seed=3, n_clusters=10000, pairs_per_cluster=100, n_chrom=8, genes_per_chrom=40)
        margs, sam = files.make_matealign_dataset(os.path.join(d, "m"), seed=4, n_pairs=300000)
        sam_path = os.path.join(d, "m", "in.sam")
        open(sam_path, "wb").write(sam)
        jobs = {"dosplitalign": (sargs + ["-a", os.path.join(d, "s", "out.tmp")], None, "1 M read pairs, 10 000 clusters"),
                "matealign": (margs, sam_path, "300 k pairs of 150 bp, search length 1000")}
        for tool, (args, stdin_path, what) in jobs.items():
            ph = phases([os.path.join(BIN, tool)] + args, stdin_path, env)
            out[tool] = {"input": what, "phase_ms_min_of_5": ph, "host_ms": round(sum(ph.values()), 1)}
            if other:
                a = digest([os.path.join(BIN, tool)] + args, stdin_path, env, os.path.join(d, "a.txt"))
                b = digest([os.path.join(other, tool)] + args, stdin_path, env, os.path.join(d, "b.txt"))
                ph_other = phases([os.path.join(other, tool)] + args, stdin_path, env)
                out[tool].update(calls=len(a.splitlines()), same_work_as_other_build=(a == b),
                                 other_build_host_ms=round(sum(ph_other.values()), 1), other_build_phase_ms_min_of_5=ph_other)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
