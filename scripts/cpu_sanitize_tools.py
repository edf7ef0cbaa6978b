#!/usr/bin/env python
"""AddressSanitizer + UBSan over the tools' host code, without a GPU: builds the five tools with
-fsanitize=address,undefined into a temporary directory and runs the CPU tests that execute tool binaries
(host logic over the device double, evalsplitalign, CLI errors) against those builds.
Usage: python scripts/cpu_sanitize_tools.py [--thread]   (exit code = pytest's; --thread = ThreadSanitizer instead)"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "defuse_b200", "host")
TOOLS = ["localalign", "matealign", "dosplitalign", "evalsplitalign", "splitseq"]


def main():
    san = "thread" if "--thread" in sys.argv[1:] else "address,undefined"
    with tempfile.TemporaryDirectory() as out:
        for t in TOOLS:
            subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=" + san, "-fno-omit-frame-pointer",
                            "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", os.path.join(out, t),
                            os.path.join(HOST, t + ".cpp"), "-L" + os.path.join(ROOT, "defuse_b200"), "-ldefuse_b200",
                            "-Wl,-rpath," + os.path.join(ROOT, "defuse_b200"), "-lpthread"], check=True)
        # the device double is preloaded in front of the sanitizer runtime: tell ASan that is intended
        os.environ["ASAN_OPTIONS"] = "verify_asan_link_order=0:detect_leaks=0:halt_on_error=1"
        os.environ["UBSAN_OPTIONS"] = "print_stacktrace=1:halt_on_error=1"
        # (error exits leave through _exit with the context thread still running, on purpose: no thread-leak reports)
        os.environ["TSAN_OPTIONS"] = "halt_on_error=1:exitcode=66:report_thread_leaks=0"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import pytest
        import test_tools_cli as tc
        import test_tools_downstream as td
        import test_tools_gpu as tg
        tg.BIN = td.BIN = tc.BIN = out
        return pytest.main(["-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] +
                           [os.path.join(ROOT, "tests", f) for f in ("test_tools_host_logic.py", "test_tools_downstream.py", "test_tools_cli.py")])


if __name__ == "__main__":
    sys.exit(main())
