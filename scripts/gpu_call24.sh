# first sweep with fewer ALU-pipe instructions (predicate-free fold, mirrored ring walked by a pointer, boundary select on
# the FMA pipe): the whole GPU suite, then the bench
TAG=${1:-r04m}
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q --timeout 240 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); e=d['e2e']; r=d['roofline']
print('value %.0f ms %.2f | sweep %.2f probe %.2f frac %.3f own %.3f step_frac %.3f | e2e %.0f ms %.2f min %.2f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['probe_sweep_ms'], r['frac'], r['frac_own_minimum'], r['step_frac_incl_probe'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0']))
for k,v in d.get('secondary',{}).items(): print('   ',k,{a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload','parity_sample')})
print(d.get('sharded_merge'))
PY
timeout 100 python scripts/gpu_fuzz.py 11 60 > gpurun_out/fuzz_$TAG.json 2> gpurun_out/fuzz_$TAG.err; echo fuzz_rc=$?; cat gpurun_out/fuzz_$TAG.json
