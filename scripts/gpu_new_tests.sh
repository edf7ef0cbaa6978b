# usage: bash scripts/gpu_new_tests.sh <tag> <pytest -k expression>
TAG=${1:-n}; EXPR=${2:-backtrace or splitseq}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$EXPR" --timeout 200 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -30 gpurun_out/pytest_$TAG.log
