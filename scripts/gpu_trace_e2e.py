#!/usr/bin/env python
"""Phase times of the end-to-end call (DFB_TRACE=1) on the bench workload: where the gap between the resident step
and the e2e step goes.  Usage: DFB_TRACE=1 python scripts/gpu_trace_e2e.py [clusters] 2> trace.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import defuse_b200 as d
import synth

clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = d.Context(0)
w = synth.split_workload(3, clusters, 100, 100, 320, 360)
host, keep = {}, []
for k in ("ref_bytes", "ref_off", "read_bytes", "read_off", "task_cluster", "task_read", "min_score"):
    host[k], t = bench.pinned(w[k])
    keep.append(t)
refs, reads = d.SeqTable(host["ref_bytes"], host["ref_off"]), d.SeqTable(host["read_bytes"], host["read_off"])
al = d.SplitReadAligner(2, -1, -2, False, 8, ctx=ctx)
for rep in range(3):
    sys.stderr.write("---- e2e call %d ----\n" % rep)
    sys.stderr.flush()
    t0 = time.perf_counter()
    r = al.align_batch(refs, reads, host["task_cluster"], host["task_read"], host["min_score"], copy=False)
    torch.cuda.synchronize()
    sys.stderr.write("---- e2e call %d: %.2f ms, %d rows ----\n" % (rep, (time.perf_counter() - t0) * 1e3, len(r.rows)))
