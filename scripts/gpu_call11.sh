# the GPU parity suite against the -DDFB_BOUNDS_CHECK build, then the regular suite, bench and e2e trace
TAG=${1:-r03n}
mkdir -p gpurun_out
( echo "DFB_LIB_PATH=gpurun_variants/libdefuse_b200_check.so python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -q"; \
  DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_check.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -q --timeout 300 --timeout-method thread 2>&1 | tail -4; \
  echo "(a recorded violation fails the fetch with 'device bounds check <code> failed')" ) > gpurun_out/bounds_check_suite_$TAG.txt 2>&1
cat gpurun_out/bounds_check_suite_$TAG.txt
timeout 700 python -m pytest tests -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$TAG.err
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py > /dev/null 2> gpurun_out/trace_e2e_$TAG.txt; echo trace_rc=$?
