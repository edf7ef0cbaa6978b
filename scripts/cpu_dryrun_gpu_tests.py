#!/usr/bin/env python
"""Dry run of `-m gpu` test LOGIC on a machine without a GPU: the Python binding is pointed at the device double
(tests/device_double, the one-call ABI entry points answered by the oracle; every other entry point a stub that fails),
so that generators, shapes and assertions of newly written parity tests can be checked before they are spent on a GPU
box.  Says nothing about the kernels -- only tests that go through the one-call forms (align_batch / backtrace_batch)
can run here; pick them with -k.
Usage: python scripts/cpu_dryrun_gpu_tests.py tests/test_zz_gpu_long_windows.py [-k expr ...]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import defuse_b200
    with tempfile.TemporaryDirectory() as d:
        obj, lib = os.path.join(d, "dp_oracle.o"), os.path.join(d, "libdouble_full.so")
        subprocess.run(["gcc", "-O2", "-fPIC", "-c", os.path.join(ROOT, "oracle", "dp_oracle.c"), "-o", obj], check=True)
        src = os.path.join(ROOT, "tests", "device_double", "device_double.cpp")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", lib, src, obj, "-lpthread"], check=True)
        have = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout.splitlines()}
        stubs = os.path.join(d, "stubs.c")
        open(stubs, "w").write("".join("int %s(void) { return 5; } /* DFB_ERR_STATE */\n" % s for s in defuse_b200.ABI_SYMBOLS if s not in have))
        subprocess.run(["gcc", "-O2", "-fPIC", "-c", stubs, "-o", stubs + ".o"], check=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", lib, src, obj, stubs + ".o", "-lpthread"], check=True)
        defuse_b200.LIB_PATH = lib
        import pytest
        return pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + sys.argv[1:])


if __name__ == "__main__":
    sys.exit(main())
