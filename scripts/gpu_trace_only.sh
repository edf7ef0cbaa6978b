mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_tr.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/smoke_tr.log; exit 1; }
DFB_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 3 > gpurun_out/trace.json 2> gpurun_out/trace.err; echo rc=$?
grep -n "pipelined" gpurun_out/trace.err | tail -20
python - <<'PY'
import json; d=json.load(open('gpurun_out/trace.json')); print(d['value'], d['e2e'])
PY
