# usage: bash scripts/gpu_tools_round.sh <tag>   -- tool-level tests + tool benches (small and at scale)
TAG=${1:-t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tools_gpu.py tests/test_tools_downstream.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -15 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/gpu_tool_scale.py > gpurun_out/tool_scale_$TAG.json 2> gpurun_out/tool_scale_$TAG.err; echo scale_rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/tool_scale_$TAG.json'))
print({k:v for k,v in d.items() if k!='trace_last_run'})
print('\n'.join(d['trace_last_run']))
PY
