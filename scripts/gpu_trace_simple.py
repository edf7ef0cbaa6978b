#!/usr/bin/env python
"""Phase times (DFB_TRACE=1: host laps + the device's side, chunk by chunk) of dfb_simple_align_batch at the matealign
shape (BASELINE.json configs[3]) and the localalign shape (configs[1]), from pinned host arrays.
Usage: DFB_TRACE=1 python scripts/gpu_trace_simple.py 2> trace.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import defuse_b200 as d
import synth

ctx = d.Context(0)
al = d.SimpleAligner(10, -5, -5, ctx=ctx)
for name, (n_refs, n_tasks, R, L) in {"matealign": (0, 500000, 1001, 150), "localalign": (10000, 1000000, 2001, 100)}.items():
    w = synth.local_workload(2, n_refs, n_tasks, R, L, own_window=(name == "matealign"))
    hp, keep = {}, []
    for k in ("ref_bytes", "ref_off", "seq_bytes", "seq_off", "task_ref", "task_seq"):
        hp[k], t = bench.pinned(w[k])
        keep.append(t)
    refs, seqs = d.SeqTable(hp["ref_bytes"], hp["ref_off"]), d.SeqTable(hp["seq_bytes"], hp["seq_off"])
    for rep in range(3):
        sys.stderr.write("---- %s call %d ----\n" % (name, rep))
        sys.stderr.flush()
        t0 = time.perf_counter()
        al.align_batch(refs, seqs, hp["task_ref"], hp["task_seq"])
        sys.stderr.write("---- %s call %d: %.2f ms (%d MB up) ----\n" % (
            name, rep, (time.perf_counter() - t0) * 1e3, sum(hp[k].nbytes for k in hp) >> 20))
