# usage: bash scripts/gpu_test_bench.sh <tag> [bench args]   (runs under gpurun; one GPU)
TAG=${1:-t}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
