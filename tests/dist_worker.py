"""Worker of tests/test_sharding.py: one rank of a world_size-N gloo job on CPU.  Each rank takes its shard of a
seeded split batch (defuse_b200.sharding), produces per-task results for it (with the CPU oracle standing in for
the GPU, which this box does not have), rank 0 gathers, merges in task order and compares with the unsharded run.
Also exercises the collectives bench.py uses for its timing (barrier, MAX / SUM all-reduce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402  (tests may use the oracle)
import synth  # noqa: E402
from defuse_b200 import sharding  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.split_workload(5, 24, 6, L=60, R_lo=90, R_hi=140, zipf=1.2)
    n = w["n_tasks"]
    cost = w["L"] * (w["ref_off"][2 * w["task_cluster"].astype(np.int64) + 2] - w["ref_off"][2 * w["task_cluster"].astype(np.int64)])
    shards, shard_of_cluster = sharding.shard_tasks(w["task_cluster"], cost, 24, world)
    mine = shards[rank]
    cnt, al = oracle.split_align_batch(w["ref_bytes"], w["ref_off"], w["read_bytes"], w["read_off"], w["task_cluster"][mine],
                                       w["task_read"][mine], w["min_score"][mine])
    best = np.zeros(len(mine), np.int32)
    pos = np.concatenate([[0], np.cumsum(cnt)])
    for k in range(len(mine)):
        if cnt[k]:
            best[k] = al[pos[k], 4]
    dist.barrier()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t[0]) == world
    c = torch.tensor([float(cost[mine].sum())], dtype=torch.float64)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    assert float(c[0]) == float(cost.sum())
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine, best, cnt), gathered, dst=0)
    if rank == 0:
        all_idx = np.concatenate([g[0] for g in gathered])
        assert np.array_equal(np.sort(all_idx), np.arange(n)), "shards must cover every task exactly once"
        for s in range(world):  # a cluster lives on exactly one shard unless it was cut for balance (-1)
            assert set(np.unique(shard_of_cluster[w["task_cluster"][gathered[s][0]]])) <= {s, -1}
        merged_best = sharding.merge_by_task(n, [(g[0], g[1]) for g in gathered])
        merged_cnt = sharding.merge_by_task(n, [(g[0], g[2]) for g in gathered])
        cnt0, al0 = oracle.split_align_batch(w["ref_bytes"], w["ref_off"], w["read_bytes"], w["read_off"], w["task_cluster"],
                                             w["task_read"], w["min_score"])
        pos0 = np.concatenate([[0], np.cumsum(cnt0)])
        best0 = np.array([al0[pos0[k], 4] if cnt0[k] else 0 for k in range(n)], np.int32)
        assert np.array_equal(merged_best, best0) and np.array_equal(merged_cnt, cnt0)
        loads = np.array([cost[g[0]].sum() for g in gathered], dtype=np.float64)
        print("OK world=%d tasks=%d load imbalance=%.3f" % (world, n, loads.max() / loads.mean()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
