"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo jobs over tests/dist_worker.py, plus unit checks of the
LPT partition."""
import os
import subprocess
import sys

import numpy as np
import pytest

from defuse_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def test_lpt_balances_skewed_clusters():
    rng = np.random.default_rng(1)
    cost = (1e6 / np.arange(1, 2001) ** 1.2).astype(np.int64) + rng.integers(0, 50, 2000)
    for n in (2, 4, 8):
        shard = sharding.assign_clusters(cost, n)
        loads = np.bincount(shard, weights=cost, minlength=n)
        assert loads.max() <= 1.34 * max(loads.mean(), cost.max())   # LPT: within 4/3 of the optimum
        assert set(shard) == set(range(n))


def test_heavy_clusters_are_cut_at_task_granularity():
    rng = np.random.default_rng(3)
    n_clusters = 200
    w = 1.0 / np.arange(1, n_clusters + 1) ** 1.2
    tc = np.sort(rng.choice(n_clusters, 20000, p=w / w.sum()))
    cost = np.full(tc.size, 68000.0)
    for n in (2, 4, 8):
        shards, shard_of_cluster = sharding.shard_tasks(tc, cost, n_clusters, n)
        loads = np.array([cost[s].sum() for s in shards])
        assert loads.max() <= 1.10 * loads.mean(), (n, loads)
        assert (shard_of_cluster == -1).sum() >= 1          # the top clusters were cut
        whole, _ = sharding.shard_tasks(tc, cost, n_clusters, n, split_heavy=False)
        assert max(cost[s].sum() for s in whole) >= loads.max()


def test_shards_partition_the_batch():
    rng = np.random.default_rng(2)
    tc = rng.integers(0, 50, 1000)
    shards, _ = sharding.shard_tasks(tc, np.ones(1000), 50, 4)
    allidx = np.concatenate(shards)
    assert np.array_equal(np.sort(allidx), np.arange(1000))
    for s in shards:
        assert np.all(np.diff(s) > 0)
    merged = sharding.merge_by_task(1000, [(s, tc[s]) for s in shards])
    assert np.array_equal(merged, tc)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_world(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(HERE, "dist_worker.py")]
    p = subprocess.run(cmd, capture_output=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert ("OK world=%d" % world) in p.stdout.decode()


def test_merge_rows_by_task_equals_stable_sort():
    """Rows of the shards back into the batch's task order: the linear counting placement gives what a stable sort of
    the concatenated parts gives (every part ordered by task, a task's rows all in one part)."""
    rng = np.random.default_rng(5)
    n_tasks = 5000
    dt = np.dtype([("task", np.int32), ("read_split", np.int32), ("col_begin", np.int64)])
    owner = rng.integers(0, 3, n_tasks)
    n_rows = rng.integers(0, 4, n_tasks)
    parts = []
    for s in range(3):
        tasks = np.repeat(np.flatnonzero(owner == s), n_rows[owner == s])
        part = np.zeros(len(tasks), dt)
        part["task"] = tasks
        part["read_split"] = rng.integers(0, 100, len(tasks))
        part["col_begin"] = np.arange(len(tasks)) + 1000 * s
        parts.append(part)
    got = sharding.merge_rows_by_task(n_tasks, parts)
    cat = np.concatenate(parts)
    want = cat[np.argsort(cat["task"], kind="stable")]
    assert got.dtype == want.dtype and np.array_equal(got, want)
    assert len(sharding.merge_rows_by_task(n_tasks, [parts[0][:0], parts[1][:0]])) == 0
