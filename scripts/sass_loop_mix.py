#!/usr/bin/env python
"""Instruction mix of the innermost loops of a kernel, from `cuobjdump -sass` (runs without a GPU).
Usage: python scripts/sass_loop_mix.py [mangled kernel name] > profiles/<tag>_sass_hot_loop.txt
Default kernel: dp_fast_kernel<8,13,SPLIT>, the first sweep of the bench workload."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "defuse_b200", "libdefuse_b200.so")
KERNEL = sys.argv[1] if len(sys.argv) > 1 else "_ZN3dfb14dp_fast_kernelILi8ELi13ELi1EEEvNS_10FastParamsE"

ALU = {"VIADDMNMX", "VIMNMX", "VIMNMX3", "LOP3", "IADD3", "ISETP", "SEL", "LEA", "SHF", "PRMT", "VIADD", "MOV", "POPC", "FLO", "IABS"}
FMA = {"IMAD", "FFMA", "HFMA2", "FMUL", "FADD", "HADD2"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in sass.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}

    def op(t):
        return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]

    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:[A-Z0-9.]+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            body = ins[addr[int(m.group(1), 16)]:i + 1]
            loops.append((len(body), int(m.group(1), 16), a, body))
    print("kernel %s: %d instructions, %d backward branches" % (KERNEL, len(ins), len(loops)))
    print("innermost loops that hold DP cell work (VIADDMNMX), smallest first:\n")
    shown = 0
    for n, lo, hi, body in sorted(loops, key=lambda x: x[0]):
        cells = sum(1 for _, t in body if "VIADDMNMX" in t)
        if cells < 20 or shown >= 3:
            continue
        c = collections.Counter(op(t) for _, t in body)
        alu = sum(v for k, v in c.items() if k in ALU)
        fma = sum(v for k, v in c.items() if k in FMA)
        print("loop 0x%04x-0x%04x: %d instructions  (alu pipe %d, fma pipe %d, other %d)" % (lo, hi, n, alu, fma, n - alu - fma))
        print("   " + ", ".join("%s %d" % kv for kv in c.most_common()))
        print()
        shown += 1


if __name__ == "__main__":
    main()
