mkdir -p gpurun_out
DFB_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/trace.json 2> gpurun_out/trace.err; echo rc=$?
tail -60 gpurun_out/trace.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/trace.json')); print(d['value'], d['e2e'])
PY
