# usage: bash scripts/gpu_quick2.sh <tag>  -- kernel parity tests + short bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print('roofline',d['roofline']['frac'],d['roofline']['kernel_ms'],d['roofline']['probe_sweep_ms'])
print('staging',d['roofline_staging'])
for k,v in d.get('secondary',{}).items(): print(k,{a:b for a,b in v.items() if a!='workload'})
PY
