"""The C restatement against the unmodified reference compiled here (oracle/_ref).  Runs wherever
oracle/_ref/libref_aligners.so exists (it is built in the build container and shipped with the snapshot)."""
import numpy as np
import pytest

import util


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_mod


def test_simple_random_vs_ref(ref):
    rng = np.random.default_rng(5)
    for scoring in [(10, -5, -5), (2, -1, -2), (1, -1, 1), (0, 0, 0), (-2, -3, -1), (7, 2, -3)]:
        refs, seqs, tr, ts = util.simple_batch(rng, 8, 60, (0, 200), (0, 120))
        import defuse_b200 as d
        rt, st = d.SeqTable.from_list(refs), d.SeqTable.from_list(seqs)
        a = ref.simple_align_batch(*scoring, rt.data, rt.off, st.data, st.off, tr, ts, impl="port")
        b = ref.simple_align_batch(*scoring, rt.data, rt.off, st.data, st.off, tr, ts, impl="ref")
        assert (a == b).all()


@pytest.mark.parametrize("params", [(2, -1, -2, False, 8), (2, -1, -2, True, 8), (2, -1, -2, False, 0), (1, -1, 1, False, 3),
                                    (5, -4, -3, False, -2), (0, 0, 0, False, 0)])
def test_split_random_vs_ref(ref, params):
    import defuse_b200 as d
    rng = np.random.default_rng(6)
    refs, reads, tc, trd = util.split_batch(rng, 10, 6, (0, 90), 0, 200, sub=0.03, indel=0.01, n_rate=0.01)
    rt, st = d.SeqTable.from_list(refs), d.SeqTable.from_list(reads)
    m, x, g, eg, ms = params
    for thr in (np.zeros(len(tc), np.int32), np.array([int(0.9 * m * len(reads[r])) for r in trd], np.int32)):
        ca, aa = ref.split_align_batch(rt.data, rt.off, st.data, st.off, tc, trd, thr, m, x, g, eg, ms, impl="port")
        cb, ab = ref.split_align_batch(rt.data, rt.off, st.data, st.off, tc, trd, thr, m, x, g, eg, ms, impl="ref")
        assert (ca == cb).all() and (aa == ab).all()


def test_reverse_complement_vs_ref(ref):
    rng = np.random.default_rng(7)
    for _ in range(50):
        s = bytes(rng.integers(32, 127, int(rng.integers(0, 60))).astype(np.uint8))
        assert ref.reverse_complement(s, impl="port") == ref.reverse_complement(s, impl="ref")


@pytest.mark.parametrize("params", [(2, -1, -2, False, 8), (2, -1, -2, True, 8), (1, -1, 1, False, 3), (5, -4, -3, False, -2),
                                    (3, -2, 0, False, 1)])
def test_backtrace_vs_ref(ref, params):
    """GetAlignments(backtrace=True): match lists of every alignment (up to a cap per task), port vs reference."""
    rng = np.random.default_rng(8)
    refs, reads, tc, trd = util.split_batch(rng, 6, 5, (0, 70), 0, 150, sub=0.04, indel=0.02, n_rate=0.01)
    m, x, g, eg, ms = params
    checked = 0
    for c, r in zip(tc, trd):
        read, ref1, ref2 = reads[r], refs[2 * c], refs[2 * c + 1]
        thr = int(0.6 * m * len(read))
        for which in range(6):
            n, hdr, m1, m2 = ref.ref_split_backtrace(read, ref1, ref2, thr, which, m, x, g, eg, ms)
            if which >= n:
                break
            p1, p2 = ref.split_backtrace(read, ref1, ref2, (hdr[0], hdr[1]), hdr[2], m, x, g, eg)
            assert p1.shape == m1.shape and (p1 == m1).all()
            assert p2.shape == m2.shape and (p2 == m2).all()
            checked += 1
    assert checked > 10


def test_random_parameters_vs_ref(ref):
    """Random scoring triples (positive gaps, negative matches, zeros), endGaps either way, minSplitScore and minScore
    below and above zero, alphabets from one letter to mixed case with N, empty strings: restatement == compiled
    reference for the split tuples and the simple score.  (The same loop ran for 21 000 rounds / 630 000 tasks while the
    oracle was being pinned; 150 rounds stay in the suite.)"""
    import defuse_b200 as d
    rng = np.random.default_rng(9)
    alphabets = [util.ACGT] + [np.frombuffer(a, np.uint8) for a in (b"ACGTN", b"AC", b"ACGTacgtN", b"A")]
    for _ in range(150):
        m, x, g = int(rng.integers(-2, 12)), int(rng.integers(-8, 3)), int(rng.integers(-8, 3))
        eg, mss = bool(rng.integers(0, 2)), int(rng.integers(-5, 20))
        alpha = alphabets[int(rng.integers(0, len(alphabets)))]
        refs, reads, tc, trd = [], [], [], []
        for c in range(6):
            refs += [util.rand_seq(rng, int(rng.integers(0, 120)), alpha), util.rand_seq(rng, int(rng.integers(0, 120)), alpha)]
            for _k in range(5):
                L = int(rng.integers(0, 70))
                joined = refs[-2] + refs[-1]
                if rng.random() < 0.6 and 0 < L <= len(joined):
                    s = int(rng.integers(0, len(joined) - L + 1))
                    read = util.mutate(rng, joined[s:s + L], 0.05, 0.02, 0.01)
                else:
                    read = util.rand_seq(rng, L, alpha)
                tc.append(c)
                trd.append(len(reads))
                reads.append(read)
        tc, trd = np.array(tc, np.int32), np.array(trd, np.int32)
        rt, st = d.SeqTable.from_list(refs), d.SeqTable.from_list(reads)
        thr = np.array([int(rng.integers(-5, max(1, m) * len(reads[r]) + 2)) for r in trd], np.int32)
        ca, aa = ref.split_align_batch(rt.data, rt.off, st.data, st.off, tc, trd, thr, m, x, g, eg, mss, impl="port")
        cb, ab = ref.split_align_batch(rt.data, rt.off, st.data, st.off, tc, trd, thr, m, x, g, eg, mss, impl="ref")
        assert (ca == cb).all() and (aa == ab).all(), (m, x, g, eg, mss)
        tr = rng.integers(0, len(refs), len(reads)).astype(np.int32)
        sa = ref.simple_align_batch(m, x, g, rt.data, rt.off, st.data, st.off, tr, trd, impl="port")
        sb = ref.simple_align_batch(m, x, g, rt.data, rt.off, st.data, st.off, tr, trd, impl="ref")
        assert (sa == sb).all(), (m, x, g)


def test_backtrace_random_parameters_vs_ref(ref):
    """Match lists of GetAlignments(backtrace=True) under random scoring triples / endGaps / thresholds and odd
    alphabets (the generator of test_random_parameters_vs_ref): restatement == compiled reference."""
    rng = np.random.default_rng(10)
    checked = 0
    for rnd in range(20):
        params, refs, reads, tc, trd, thr = util.random_parameter_round(rng, rnd if rnd % 4 != 3 else rnd + 1)  # short sequences only
        m, x, g, eg, mss = params
        for c, r, t in zip(tc, trd, thr):
            read, ref1, ref2 = reads[r], refs[2 * c], refs[2 * c + 1]
            for which in range(3):
                n, hdr, m1, m2 = ref.ref_split_backtrace(read, ref1, ref2, int(t), which, m, x, g, eg, mss)
                if which >= n:
                    break
                p1, p2 = ref.split_backtrace(read, ref1, ref2, (hdr[0], hdr[1]), hdr[2], m, x, g, eg)
                assert p1.shape == m1.shape and (p1 == m1).all(), (params, which)
                assert p2.shape == m2.shape and (p2 == m2).all(), (params, which)
                checked += 1
    assert checked > 100
