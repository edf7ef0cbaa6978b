"""Synthetic RNA-seq style workloads for bench.py and the tests: there is no
network, so reads are drawn from random window pairs with a planted fusion junction.
Vectorised numpy; seeded; no reference code involved."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(rng, reads, sub, n_rate):
    """In-place substitutions and N's on a (n, L) uint8 matrix."""
    if sub > 0:
        mask = rng.random(reads.shape) < sub
        reads[mask] = ACGT[rng.integers(0, 4, int(mask.sum()))]
    if n_rate > 0:
        mask = rng.random(reads.shape) < n_rate
        reads[mask] = ord("N")


def _indels(rng, src, L, rate):
    """Per-base indels on an (n, L + extra) matrix of source bases: each source position is deleted with probability
    rate/2 or gets a random base inserted in front of it with probability rate/2; the first L emitted bases are kept."""
    n, Ls = src.shape
    r = rng.random((n, Ls))
    emit = np.ones((n, Ls), dtype=np.int64)
    emit[r < rate / 2] = 0
    emit[(r >= rate / 2) & (r < rate)] = 2
    end = np.cumsum(emit, axis=1)                 # one past the slot of the source base itself
    out = np.empty((n, L), dtype=np.uint8)
    out[:] = ACGT[rng.integers(0, 4, (n, L))]     # (slots never written: only if more than `extra` bases were deleted)
    t, i = np.nonzero((emit >= 1) & (end <= L))
    out[t, end[t, i] - 1] = src[t, i]
    t, i = np.nonzero((emit == 2) & (end - 1 <= L) & (end >= 2))
    out[t, end[t, i] - 2] = ACGT[rng.integers(0, 4, t.size)]
    return out


STRESS = dict(L=250, R_lo=700, R_hi=860, sub=0.08, n_rate=0.01, indel_frac=0.0, indel_rate=0.02, zipf=1.2,
              polya_frac=0.05, lower_frac=0.01, polya_window_frac=0.01)
"""SURVEY 8(d) config 5 as written: 250-bp reads, windows of a 600 +- 80 bp library, 8 % substitutions, 2 % indels, 1 % N,
5 % poly-A-tail reads, lowercase runs in 1 % of the windows, Zipf(1.2) cluster sizes; in 1 % of the clusters the junction
lies inside a poly-A run (deFuse appends 50 A's to every cDNA, NEWS.md:81), which is what makes spanning reads tie-heavy."""


def split_workload(seed, n_clusters, tasks_per_cluster, L=100, R_lo=320, R_hi=360, sub=0.01, n_rate=0.001,
                   indel_frac=0.1, span_frac=0.5, unrelated_frac=0.05, zipf=None, match=2, indel_rate=0.0,
                   polya_frac=0.0, lower_frac=0.0, polya_window_frac=0.0):
    """dosplitalign-shaped batch (configs 1/3/5 of BASELINE.json): n_clusters window pairs of R_lo..R_hi
    bases with a planted junction; per task a read of L bases that spans the junction (span_frac), lies
    wholly in one window, or is unrelated.  Returns CSR tables + task arrays + minScore per task."""
    rng = np.random.default_rng(seed)
    R1 = rng.integers(R_lo, R_hi + 1, n_clusters)
    R2 = rng.integers(R_lo, R_hi + 1, n_clusters)
    Rmax = int(R_hi)
    win1 = ACGT[rng.integers(0, 4, (n_clusters, Rmax))]
    win2 = ACGT[rng.integers(0, 4, (n_clusters, Rmax))]
    bp1 = rng.integers(R1 // 3, R1 - 4)          # fusion = win1[:bp1] + win2[bp2:R2]
    bp2 = rng.integers(4, 2 * R2 // 3)
    if polya_window_frac > 0:
        # the junction of some clusters lies inside a poly-A run (30 A's on either side): the split point of a spanning
        # read can sit anywhere in the run, so dozens of split rows tie (the reference enumerates them all)
        for c in np.nonzero(rng.random(n_clusters) < polya_window_frac)[0]:
            win1[c, max(0, bp1[c] - 30):bp1[c]] = ord("A")
            win2[c, bp2[c]:bp2[c] + 30] = ord("A")
    # fusion sequence per cluster, padded
    flen = bp1 + (R2 - bp2)
    fus = np.zeros((n_clusters, 2 * Rmax), dtype=np.uint8)
    col = np.arange(2 * Rmax)[None, :]
    left = col < bp1[:, None]
    fus[:, :Rmax][left[:, :Rmax]] = win1[left[:, :Rmax]]
    src2 = col - bp1[:, None] + bp2[:, None]
    right = (~left) & (src2 < R2[:, None])
    fus[right] = win2[np.nonzero(right)[0], src2[right]]

    if zipf:
        w = 1.0 / np.arange(1, n_clusters + 1) ** zipf
        w /= w.sum()
        n_tasks = n_clusters * tasks_per_cluster
        task_cluster = np.sort(rng.choice(n_clusters, n_tasks, p=w)).astype(np.int32)
    else:
        task_cluster = np.repeat(np.arange(n_clusters, dtype=np.int32), tasks_per_cluster)
    n_tasks = task_cluster.size
    c = task_cluster
    kind = rng.random(n_tasks)
    # start positions in the fusion sequence
    lo_span = np.maximum(0, bp1[c] - L + 4)
    hi_span = np.maximum(lo_span, np.minimum(bp1[c] - 4, flen[c] - L))
    start_span = lo_span + (rng.random(n_tasks) * (hi_span - lo_span + 1)).astype(np.int64)
    start_any = (rng.random(n_tasks) * np.maximum(1, flen[c] - L + 1)).astype(np.int64)
    start = np.where(kind < span_frac, start_span, start_any)
    extra = 24 if indel_rate > 0 else 0
    idx = start[:, None] + np.arange(L + extra)[None, :]
    idx = np.minimum(idx, 2 * Rmax - 1)
    reads = fus[c[:, None], idx]
    unrelated = kind > 1.0 - unrelated_frac
    reads[unrelated] = ACGT[rng.integers(0, 4, (int(unrelated.sum()), L + extra))]
    if indel_rate > 0:
        reads = _indels(rng, reads, L, indel_rate)
    if polya_frac > 0:
        # poly-A tails: the last 30..80 bases of some reads (before the substitutions, so the tails are not perfect)
        for t in np.nonzero(rng.random(n_tasks) < polya_frac)[0]:
            reads[t, L - int(rng.integers(30, min(81, L))):] = ord("A")
    _mutate(rng, reads, sub, n_rate)
    # crude single-base deletions: shift the tail left and append a random base
    dele = np.nonzero(rng.random(n_tasks) < indel_frac)[0]
    if dele.size:
        pos = rng.integers(1, L - 1, dele.size)
        for t, p in zip(dele[:20000], pos[:20000]):
            reads[t, p:-1] = reads[t, p + 1:]
            reads[t, -1] = ACGT[rng.integers(0, 4)]

    ref_len = np.empty(2 * n_clusters, dtype=np.int64)
    ref_len[0::2] = R1
    ref_len[1::2] = R2
    ref_off = np.zeros(2 * n_clusters + 1, dtype=np.int64)
    ref_off[1:] = np.cumsum(ref_len)
    ref_bytes = np.zeros(int(ref_off[-1]), dtype=np.uint8)
    if lower_frac > 0:
        # soft-masked (lowercase) runs of 20..80 bases in some windows: bytes compare case-sensitively
        for win, R in ((win1, R1), (win2, R2)):
            for k in np.nonzero(rng.random(n_clusters) < lower_frac)[0]:
                a = int(rng.integers(0, max(1, R[k] - 80)))
                win[k, a:a + int(rng.integers(20, 81))] |= 0x20
    both = np.stack([win1, win2], axis=1).reshape(2 * n_clusters, Rmax)
    keep = np.arange(Rmax)[None, :] < ref_len[:, None]
    ref_bytes[:] = both[keep]
    read_off = np.arange(n_tasks + 1, dtype=np.int64) * L
    task_read = np.arange(n_tasks, dtype=np.int32)
    min_score = np.full(n_tasks, int(np.float64(np.float32(np.float32(L) * np.float32(match))) * 0.90), dtype=np.int32)
    cells = int((L * (R1[c] + R2[c])).sum())
    return {"ref_bytes": ref_bytes, "ref_off": ref_off, "read_bytes": np.ascontiguousarray(reads).reshape(-1),
            "read_off": read_off, "task_cluster": task_cluster, "task_read": task_read, "min_score": min_score,
            "cells": cells, "n_tasks": int(n_tasks), "L": int(L)}


def local_workload(seed, n_refs, n_tasks, R=2001, L=100, sub=0.02, unrelated_frac=0.2, own_window=False):
    """localalign-shaped batch (config 2): n_refs references of R bases; each task picks one and a
    L-base substring with substitutions, or an unrelated random sequence.  own_window: every task has its own
    reference, in task order -- what matealign submits (one window cut per discordant mate, matealign.cpp: task_ref[k] = k)."""
    rng = np.random.default_rng(seed)
    if own_window:
        n_refs = n_tasks
    refs = ACGT[rng.integers(0, 4, (n_refs, R))]
    task_ref = np.arange(n_tasks, dtype=np.int32) if own_window else rng.integers(0, n_refs, n_tasks).astype(np.int32)
    start = rng.integers(0, R - L + 1, n_tasks)
    seqs = refs[task_ref[:, None], start[:, None] + np.arange(L)[None, :]]
    unrelated = rng.random(n_tasks) < unrelated_frac
    seqs[unrelated] = ACGT[rng.integers(0, 4, (int(unrelated.sum()), L))]
    _mutate(rng, seqs, sub, 0.0)
    return {"ref_bytes": np.ascontiguousarray(refs).reshape(-1), "ref_off": np.arange(n_refs + 1, dtype=np.int64) * R,
            "seq_bytes": np.ascontiguousarray(seqs).reshape(-1), "seq_off": np.arange(n_tasks + 1, dtype=np.int64) * L,
            "task_ref": task_ref, "task_seq": np.arange(n_tasks, dtype=np.int32),
            "cells": int(n_tasks) * R * L, "n_tasks": int(n_tasks), "L": int(L)}
