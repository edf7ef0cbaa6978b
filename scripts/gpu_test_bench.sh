# usage: bash scripts/gpu_test_bench.sh <tag> [bench args]   (runs under gpurun; one GPU)
# every step is bounded: a hung kernel must not burn the GPU budget
TAG=${1:-t}; shift
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 420 python -m pytest tests -m gpu -x -q --timeout 90 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
