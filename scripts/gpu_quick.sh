# quick perf check: smoke (parity gate) + bench without extras.  usage: bash scripts/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1 || { echo SMOKE_FAILED; tail -20 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
python - <<PY
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'], 'sweep', d['roofline']['kernel_ms'], 'probe', d['roofline']['probe_sweep_ms'], 'frac', d['roofline']['frac'])
PY
