# usage: bash scripts/gpu_sanitize.sh <memcheck|racecheck> <tag>
# ONE compute-sanitizer tool per gpurun call (B200_PROFILING.md), after the same tests have passed without it, over the
# small parity tests that exercise every kernel family: sweeps (simple / split), probe rounds, assembly, job build
# (the pipelined test forces the device-built path), pack, backtrace.
TOOL=${1:-memcheck}; TAG=${2:-s}
mkdir -p gpurun_out
SEL="planted or edge_cases or simple_random or pipelined or event_buffer or staged or backtrace_dosplitalign"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 --timeout-method thread -k "$SEL" > gpurun_out/sanitize_plain_$TAG.log 2>&1 || { echo PLAIN_RUN_FAILED; tail -5 gpurun_out/sanitize_plain_$TAG.log; exit 1; }
tail -1 gpurun_out/sanitize_plain_$TAG.log
timeout 2400 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 1800 --timeout-method thread \
  -k "$SEL" > gpurun_out/sanitize_${TOOL}_$TAG.log 2>&1
echo sanitize_rc=$?
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|out of bounds|hazard" gpurun_out/sanitize_${TOOL}_$TAG.log | head -20
tail -3 gpurun_out/sanitize_${TOOL}_$TAG.log
