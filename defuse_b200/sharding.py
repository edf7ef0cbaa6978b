"""Partitioning of a batch across GPUs: by candidate cluster, no exchange between shards.

Every (read x cluster) task is independent, so a shard is just a subset of tasks; keeping a cluster on one GPU
means its window pair is uploaded once.  Clusters are dealt out longest-processing-time first on their DP cell
count; results come back per task and are merged in the original task order (the reference's emission order)."""
import numpy as np


def assign_clusters(cluster_cost, n_shards):
    """LPT: heaviest cluster first onto the currently lightest shard.  Returns shard id per cluster (int32)."""
    cluster_cost = np.asarray(cluster_cost, dtype=np.float64)
    shard = np.zeros(cluster_cost.size, dtype=np.int32)
    load = np.zeros(n_shards, dtype=np.float64)
    for c in np.argsort(-cluster_cost, kind="stable"):
        s = int(np.argmin(load))
        shard[c] = s
        load[s] += cluster_cost[c]
    return shard


def shard_tasks(task_cluster, task_cost, n_clusters, n_shards, split_heavy=True):
    """Task indices of every shard (ascending inside a shard, so a shard keeps the batch's task order).

    Whole clusters are the unit, except clusters heavier than a quarter of the mean shard load: those are cut into
    runs of their tasks (any GPU can run any task; the cluster's window pair is then uploaded to each GPU that
    holds a run).  Returns (list of index arrays, shard id per cluster or -1 for clusters that were cut)."""
    task_cluster = np.asarray(task_cluster, dtype=np.int64)
    task_cost = np.asarray(task_cost, dtype=np.float64)
    n_tasks = task_cluster.size
    cost = np.bincount(task_cluster, weights=task_cost, minlength=n_clusters)
    cap = max(1.0, cost.sum() / max(1, n_shards) / 4.0)
    # units: (first task position in cluster-sorted order, count, cost)
    order = np.argsort(task_cluster, kind="stable")
    starts = np.concatenate([[0], np.cumsum(np.bincount(task_cluster, minlength=n_clusters))])
    unit_lo, unit_hi, unit_cost, unit_cluster = [], [], [], []
    for c in range(n_clusters):
        lo, hi = int(starts[c]), int(starts[c + 1])
        if hi == lo:
            continue
        if split_heavy and n_shards > 1 and cost[c] > cap:
            csum = np.cumsum(task_cost[order[lo:hi]])
            pieces = int(np.ceil(cost[c] / cap))
            cuts = np.unique(np.searchsorted(csum, np.linspace(0, csum[-1], pieces + 1)[1:-1]))
            bounds = [lo] + [lo + int(k) for k in cuts if 0 < k < hi - lo] + [hi]
        else:
            bounds = [lo, hi]
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b > a:
                unit_lo.append(a)
                unit_hi.append(b)
                unit_cost.append(task_cost[order[a:b]].sum())
                unit_cluster.append(c)
    shard_of_unit = assign_clusters(np.asarray(unit_cost), n_shards)
    shard_of_task = np.zeros(n_tasks, dtype=np.int32)
    shard_of_cluster = np.full(n_clusters, -2, dtype=np.int32)
    for u in range(len(unit_lo)):
        shard_of_task[order[unit_lo[u]:unit_hi[u]]] = shard_of_unit[u]
        c = unit_cluster[u]
        shard_of_cluster[c] = shard_of_unit[u] if shard_of_cluster[c] in (-2, shard_of_unit[u]) else -1
    return [np.nonzero(shard_of_task == s)[0] for s in range(n_shards)], shard_of_cluster


def merge_by_task(n_tasks, parts):
    """parts: iterable of (task_indices, values) from the shards -> one array in task order."""
    out = None
    for idx, val in parts:
        val = np.asarray(val)
        if out is None:
            out = np.zeros((n_tasks,) + val.shape[1:], dtype=val.dtype)
        out[np.asarray(idx)] = val
    return out


def merge_rows_by_task(n_tasks, row_parts):
    """Row records of the shards (structured arrays with a `task` field in the BATCH's numbering, every part ordered by
    task -- the order dfb_split_align_batch returns) -> one array in task order, the rows of a task in their part's
    order.  Counting placement: rows per task, an exclusive scan, every part copied to its slots -- linear in the number
    of rows, no sort (a task's rows all come from one part)."""
    parts = [np.asarray(p) for p in row_parts if len(p)]
    if not parts:
        return np.zeros(0, dtype=np.asarray(row_parts[0]).dtype if len(row_parts) else np.int32)
    counts = np.zeros(n_tasks + 1, dtype=np.int64)
    for p in parts:
        counts[1:] += np.bincount(p["task"], minlength=n_tasks)
    first = np.cumsum(counts)  # first[t] = slot of task t's first row
    out = np.empty(int(first[-1]), dtype=parts[0].dtype)
    for p in parts:
        task = p["task"].astype(np.int64)
        # position of a row inside its task's run: its index minus the index of the run's first row
        run_start = np.flatnonzero(np.concatenate([[True], task[1:] != task[:-1]]))
        run_len = np.diff(np.concatenate([run_start, [len(task)]]))
        within = np.arange(len(task)) - np.repeat(run_start, run_len)
        # (whole records moved as opaque items: assignment field by field through a structured dtype is 7x slower)
        opaque = np.dtype((np.void, out.dtype.itemsize))
        out.view(opaque)[first[task] + within] = np.ascontiguousarray(p).view(opaque)
    return out
