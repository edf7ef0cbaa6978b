mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke_rc=$?
tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_rc=$?
tail -30 gpurun_out/pytest_gpu.log
python - > gpurun_out/micro.log 2>&1 <<'PY'
import defuse_b200 as d
ctx = d.default_context(0)
info = ctx.device_info(); print(info)
for k, name in d.MICROBENCH_KINDS.items():
    r, ms = ctx.microbench_issue_rate(k, 2000)
    print("%-22s %8.3f Gwarp-instr/s  %.3f ms  -> %.2f warp-instr/clk/SM @%.0f MHz" % (name, r/1e9, ms, r/info['sm_count']/(info['clock_khz']*1e3), info['clock_khz']/1e3))
PY
cat gpurun_out/micro.log
