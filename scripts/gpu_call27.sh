# ring decode with one PRMT per column (both halves): kernel parity, then the bench
TAG=${1:-r04s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 240 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -2 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); e=d['e2e']; r=d['roofline']
print('value %.0f ms %.3f | sweep %.3f probe %.3f frac %.3f own %.3f | e2e %.0f ms %.2f min %.2f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['probe_sweep_ms'], r['frac'], r['frac_own_minimum'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0']))
for k,v in d.get('secondary',{}).items(): print('   ',k,{a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload','parity_sample')})
PY
