# one full GPU pass without profilers: smoke, tests, bench (+reference arm), tool-level bench
# usage: bash scripts/gpu_round_full.sh <tag>
TAG=${1:-r}
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/benchref_$TAG.json 2> gpurun_out/benchref_$TAG.err; echo benchref_rc=$?
# (profilers run in their own gpurun calls: scripts/gpu_ncu.sh for ncu, scripts/gpu_sanitize.sh <tool> for compute-sanitizer)
timeout 600 python scripts/gpu_tools_bench.py > gpurun_out/tools_bench_$TAG.json 2> gpurun_out/tools_bench_$TAG.err; echo tools_rc=$?
cat gpurun_out/tools_bench_$TAG.json; tail -3 gpurun_out/tools_bench_$TAG.err
timeout 300 python scripts/gpu_tool_scale.py > gpurun_out/tool_scale_split_$TAG.json 2> gpurun_out/tool_scale_split_$TAG.err; echo scale_split_rc=$?
timeout 300 python scripts/gpu_tool_scale.py localalign 1000000 10000 > gpurun_out/tool_scale_local_$TAG.json 2> gpurun_out/tool_scale_local_$TAG.err; echo scale_local_rc=$?
timeout 300 python scripts/gpu_tool_scale.py matealign 300000 > gpurun_out/tool_scale_mate_$TAG.json 2> gpurun_out/tool_scale_mate_$TAG.err; echo scale_mate_rc=$?
timeout 120 python scripts/gpu_fuzz.py 7 60 > gpurun_out/fuzz_$TAG.json 2> gpurun_out/fuzz_$TAG.err; echo fuzz_rc=$?; cat gpurun_out/fuzz_$TAG.json
