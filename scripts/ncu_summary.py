#!/usr/bin/env python
"""Turns the `ncu --set full` capture of the two sweeps (scripts/gpu_ncu_full.sh, brought back in gpurun_out/) into the
committed evidence: profiles/<tag>_ncu_full_sweep_and_probe.csv (raw page) and profiles/ncu_traffic.json (the figures
bench.py quotes: DRAM bytes per task of the first sweep, its ALU-pipe share).  Runs in the build container (ncu -i).
Usage: python scripts/ncu_summary.py <tag> [tasks per launch = 200000]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unit_scale(unit):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0,
            "s": 1e3, "second": 1e3, "ns": 1e-6, "nsecond": 1e-6}.get(unit, 1.0)


def main():
    tag = sys.argv[1]
    tasks = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
    out_csv = os.path.join(ROOT, "profiles", "%s_ncu_full_sweep_and_probe.csv" % tag)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(out_csv, "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        return v * unit_scale(units[col[name]])

    summary = {"source": "profiles/%s (ncu --set full --clock-control none, bench.py --clusters 2000: one %d-task launch per kernel)"
                         % (os.path.basename(out_csv), tasks), "capture": tag}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "dp_fast_kernel_split" if "dp_fast_kernel" in name else ("dp_probe_kernel" if "dp_probe_kernel" in name else None)
        if not key:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        summary[key + "_ms"] = val(r, "gpu__time_duration.sum")
        summary[key + "_dram_bytes_read"] = rd
        summary[key + "_dram_bytes_written"] = wr
        summary[key + "_instructions"] = val(r, "smsp__inst_executed.sum")
        summary[key + "_alu_pipe_pct"] = round(val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2)
        summary[key + "_fma_pipe_pct"] = round(val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2) \
            if "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active" in col else None
        summary[key + "_issue_active_pct"] = round(val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), 2)
        summary[key + "_shared_bank_conflicts"] = val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        summary[key + "_registers"] = int(val(r, "launch__registers_per_thread"))
        if key == "dp_fast_kernel_split":
            summary["dp_fast_kernel_split_dram_bytes_per_task"] = (rd + wr) / tasks
            summary["dp_fast_kernel_split_note"] = ("the writes are the wavefront checkpoints (one every 4G steps) that let the probe "
                                                    "sweep resume block by block instead of re-sweeping")
    json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
