"""Split alignment against window pairs far longer than deFuse ever cuts (hundreds of bases): beyond what the probe sweep
of the s16x2 kernels can address by checkpoint block (255 blocks of 4G steps: 8 160 / 16 320 / 32 640 columns for the
G = 8 / 16 / 32 classes) and beyond the 16-bit reference length of a job pair (65 535).  Such tasks are swept by the s32
kernels; the junctions are planted near the END of the windows, where a truncated sweep would miss the arg-max columns.
(Sorted last on purpose: the path is new in the suite.)"""
import numpy as np
import pytest

import util
from test_gpu_parity import _check_split

pytestmark = pytest.mark.gpu


def _late_junction_batch(rng, R1, R2, L, n_reads):
    """One cluster whose reads span a junction in the last few hundred columns of window 1 / first columns of the
    reversed window 2, plus reads from either side and random ones."""
    ref1, ref2 = util.rand_seq(rng, R1), util.rand_seq(rng, R2)
    bp1 = R1 - int(rng.integers(5, 200))
    bp2 = int(rng.integers(5, 200))
    fusion = ref1[:bp1] + ref2[bp2:]
    reads = []
    for k in range(n_reads):
        kind = k % 4
        if kind == 0:
            s = bp1 - int(rng.integers(10, L - 10))
            read = fusion[s:s + L]
        elif kind == 1:
            s = int(rng.integers(max(0, R1 - 3 * L), R1 - L))
            read = ref1[s:s + L]
        elif kind == 2:
            s = int(rng.integers(0, min(R2 - L, 3 * L)))
            read = ref2[s:s + L]
        else:
            read = util.rand_seq(rng, L)
        reads.append(util.mutate(rng, read, 0.01, 0.002))
    return [ref1, ref2], reads


@pytest.mark.parametrize("R,L,mixed", [(8100, 100, False), (8300, 100, False), (9000, 100, False), (20000, 100, False),
                                       (17000, 300, False), (34000, 700, False), (70000, 100, False), (9000, 100, True)])
def test_split_windows_longer_than_the_probe_can_address(oracle_mod, gpu_ctx, R, L, mixed):
    import defuse_b200 as d
    rng = np.random.default_rng(R + L)
    refs, reads = _late_junction_batch(rng, R, R - 37, L, 8)
    task_cluster = [0] * len(reads)
    if mixed:
        # an ordinary cluster in the same batch: both kernel families run side by side (last case of the list)
        refs2, reads2, _, _ = util.split_batch(rng, 1, 6, L, 300, 380)
        refs, reads, task_cluster = refs + refs2, reads + reads2, task_cluster + [1] * len(reads2)
    task_cluster = np.array(task_cluster, np.int32)
    task_read = np.arange(len(reads), dtype=np.int32)
    min_score = np.array([d.split_min_score(len(r)) for r in reads], np.int32)
    res = _check_split(oracle_mod, gpu_ctx, refs, reads, task_cluster, task_read, min_score)
    assert (res.best[:8] > 0).sum() >= 2  # the planted junction near the window end is found


@pytest.mark.parametrize("L", [330, 400, 512, 640, 768, 900, 1024])
def test_split_long_reads_in_the_wide_classes(oracle_mod, gpu_ctx, L):
    """Split mode of the (16,25) (16,32) (32,24) (32,32) classes of the s16x2 kernel (reads of 321-1024 bases; deFuse's
    reads are 50-250): first sweep with checkpoints every 64 / 128 steps, probe sweep, assembly."""
    import defuse_b200 as d
    rng = np.random.default_rng(L)
    refs, reads, tc, trd = util.split_batch(rng, 3, 4, (L - 6, L), L + 150, L + 700, sub=0.02, indel=0.004, n_rate=0.003)
    ms = np.array([d.split_min_score(len(reads[r])) for r in trd], np.int32)
    res = _check_split(oracle_mod, gpu_ctx, refs, reads, tc, trd, ms)
    assert (res.best > 0).sum() >= 3


@pytest.mark.parametrize("R,L", [(30000, 100), (65535, 150), (65536, 150), (100000, 600)])
def test_simple_very_long_references(oracle_mod, gpu_ctx, R, L):
    """SimpleAligner against references at and beyond the 16-bit length of a job pair (s16x2 up to 65 535, s32 beyond),
    the related reads taken from the far end of the reference."""
    from test_gpu_parity import _check_simple
    rng = np.random.default_rng(R + L)
    refs = [util.rand_seq(rng, R), util.rand_seq(rng, R - 1)]
    seqs, task_ref = [], []
    for k in range(10):
        r = k % 2
        if k % 5 == 4:
            seq = util.rand_seq(rng, L)
        else:
            s = len(refs[r]) - L - int(rng.integers(0, 300))
            seq = util.mutate(rng, refs[r][s:s + L], 0.03, 0.004)
        seqs.append(seq)
        task_ref.append(r)
    _check_simple(oracle_mod, gpu_ctx, refs, seqs, np.array(task_ref, np.int32), np.arange(len(seqs), dtype=np.int32), 10, -5, -5)
