"""The ctypes mirror's own logic (result views, buffer reuse, expansion of the factorised rows) without a GPU: the binding is
pointed at the device double (tests/device_double: the one-call ABI entry points answered by the oracle) in a child
process.  Says nothing about the kernels -- the `-m gpu` parity tests do that."""
import os
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CHILD = r'''
import os, subprocess, sys, tempfile
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import defuse_b200 as d
import oracle, util
with tempfile.TemporaryDirectory() as tmp:
    obj, lib = os.path.join(tmp, "dp_oracle.o"), os.path.join(tmp, "libdouble_full.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-c", os.path.join(ROOT, "oracle", "dp_oracle.c"), "-o", obj], check=True)
    src = os.path.join(ROOT, "tests", "device_double", "device_double.cpp")
    inc = "-I" + os.path.join(ROOT, "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", inc, "-o", lib, src, obj, "-lpthread"], check=True)
    have = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout.splitlines()}
    stubs = os.path.join(tmp, "stubs.c")
    open(stubs, "w").write("".join("int %s(void) { return 5; }\n" % s for s in d.ABI_SYMBOLS if s not in have))
    subprocess.run(["gcc", "-O2", "-fPIC", "-c", stubs, "-o", stubs + ".o"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", inc, "-o", lib, src, obj, stubs + ".o", "-lpthread"], check=True)
    d.LIB_PATH = lib
    ctx = d.Context(0)
    rng = np.random.default_rng(3)
    refs, reads, tc, tr = util.split_batch(rng, 6, 5, (20, 80), 60, 200, sub=0.02)
    rt, st = d.SeqTable.from_list(refs), d.SeqTable.from_list(reads)
    ms = np.array([d.split_min_score(len(reads[r])) for r in tr], np.int32)
    al = d.SplitReadAligner(ctx=ctx)
    cnt, want = oracle.split_align_batch(rt.data, rt.off, st.data, st.off, tc, tr, ms)
    pos = np.concatenate([[0], np.cumsum(cnt)])

    def check(res):
        for t in range(len(tc)):
            a, w = res.alignments(t), want[pos[t]:pos[t + 1]]
            assert a.shape == w.shape and (a == w).all(), t

    # copy=True: arrays of the caller's own, untouched by later calls
    r1 = al.align_batch(rt, st, tc, tr, ms)
    best1 = r1.best.copy()
    check(r1)
    r2 = al.align_batch(rt, st, tc[:7], tr[:7], ms[:7])
    assert r1.best.base is None or r1.best.base is not r2.best.base
    assert (r1.best == best1).all() and len(r2.best) == 7 and (r2.best == best1[:7]).all()
    check(r1)
    # copy=False: views, `best` in a buffer of the aligner that the next call reuses
    v1 = al.align_batch(rt, st, tc, tr, ms, copy=False)
    check(v1)
    assert (v1.best == best1).all()
    v2 = al.align_batch(rt, st, tc[::-1].copy(), tr[::-1].copy(), ms[::-1].copy(), copy=False)
    assert np.shares_memory(v1.best, v2.best)
    assert (v2.best == best1[::-1]).all()
    # a larger batch grows the buffer, a smaller one is a prefix view of it
    tc3, tr3, ms3 = np.concatenate([tc, tc]), np.concatenate([tr, tr]), np.concatenate([ms, ms])
    v3 = al.align_batch(rt, st, tc3, tr3, ms3, copy=False)
    assert len(v3.best) == 2 * len(tc) and (v3.best[:len(tc)] == best1).all() and (v3.best[len(tc):] == best1).all()
    v4 = al.align_batch(rt, st, tc[:3], tr[:3], ms[:3], copy=False)
    assert len(v4.best) == 3 and np.shares_memory(v3.best, v4.best)
    # SimpleAligner: the output array is filled completely, caller-provided or not
    refs2, seqs2, t_ref, t_seq = util.simple_batch(rng, 5, 40, (1, 120), (0, 60))
    rt2, st2 = d.SeqTable.from_list(refs2), d.SeqTable.from_list(seqs2)
    sa = d.SimpleAligner(10, -5, -5, ctx=ctx)
    want2 = oracle.simple_align_batch(10, -5, -5, rt2.data, rt2.off, st2.data, st2.off, t_ref, t_seq)
    assert (sa.align_batch(rt2, st2, t_ref, t_seq) == want2).all()
    out = np.full(len(t_ref), -7, np.int32)
    assert sa.align_batch(rt2, st2, t_ref, t_seq, out=out) is out and (out == want2).all()
print("binding ok")
'''


def test_binding_over_device_double(tmp_path):
    script = tmp_path / "child.py"
    script.write_text(textwrap.dedent(CHILD))
    p = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "binding ok" in p.stdout
