# compute-sanitizer memcheck over the small parity tests that exercise every kernel family
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 --timeout-method thread \
  -k "planted or varied or ties or edge_cases or generic_params or event_buffer or pipelined or backtrace or simple_random or staged" > gpurun_out/sanitize_${1:-s}.log 2>&1
echo sanitize_rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_${1:-s}.log | head -20
tail -3 gpurun_out/sanitize_${1:-s}.log
