# parity suite, bench, e2e phase trace; then ncu of the sweeps (small config)
TAG=${1:-r03c}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -30 gpurun_out/smoke_$TAG.log; }
timeout 700 python -m pytest tests -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$TAG.err
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py > /dev/null 2> gpurun_out/trace_e2e_$TAG.txt; echo trace_rc=$?
SMALL="python bench.py --steps 2 --warmup 1 --clusters 2000 --no-cpu-baseline --no-secondary --e2e-steps 1"
timeout 200 $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
echo ncu_launch_rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_(fast|probe)_kernel' -s 2 -c 2 -o gpurun_out/prof_$TAG $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo ncu_full_rc=$?
tail -3 gpurun_out/ncu_full_$TAG.log
