# last sanity pass of the committed build: one-call parity tests, the default bench invocation, the reference arm
TAG=${1:-r04r}
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo smoke_rc=$?; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method thread \
  -k "pipelined or overflow or large or staged or planted or edge" > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -2 gpurun_out/pytest_$TAG.log
timeout 500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); e=d['e2e']; r=d['roofline']
print('value %.0f ms %.2f | sweep %.2f probe %.2f frac %.3f own %.3f | e2e %.0f ms %.2f min %.2f | alu %s capture %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['probe_sweep_ms'], r['frac'], r['frac_own_minimum'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'], r['ncu_alu_pipe_pct'], r['ncu_capture']))
print('cpu', d['cpu_baseline']['value'], 'gpu_launches', d['gpu_launches'], 'traffic', r['traffic'])
PY
