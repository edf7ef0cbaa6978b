# parity suite, bench, e2e phase trace (two-stage lane 1, pipelined simple batches)
TAG=${1:-r03g}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -30 gpurun_out/smoke_$TAG.log; }
timeout 700 python -m pytest tests -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$TAG.err
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py > /dev/null 2> gpurun_out/trace_e2e_$TAG.txt; echo trace_rc=$?
