# host path of the one-call forms: polled event waits, copy-only upload streams (one in order, or two alternating),
# large pinned uploads over two copy engines, more chunks for simple batches with references in task order.
# A/B by environment knobs inside one call; parity of the pipelined paths first.
TAG=${1:-r04d}
mkdir -p gpurun_out
python scripts/gpu_h2d_rate.py > gpurun_out/h2d_rate_$TAG.txt 2>&1; cat gpurun_out/h2d_rate_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method thread \
  -k "pipelined or overflow or full_size or large or staged" > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo ${name}_rc=$?
}
run old DFB_WAIT=block DFB_UPLOAD_STREAMS=2 DFB_H2D_SPLIT=0 DFB_SIMPLE_CHUNK_TASKS=150000
run new DFB_X=1
run new_waitblock DFB_WAIT=block
run new_streams2 DFB_UPLOAD_STREAMS=2
run new_nosplit DFB_H2D_SPLIT=0
run old_dev DFB_DEVICE_BUILD=1 DFB_WAIT=block DFB_UPLOAD_STREAMS=2 DFB_H2D_SPLIT=0 DFB_SIMPLE_CHUNK_TASKS=150000
run new_dev DFB_DEVICE_BUILD=1
run new2 DFB_X=1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']; r=d['roofline']; s=d.get('secondary',{})
        print('%-16s value %.0f ms %.2f | e2e %.0f ms %.2f min %.2f med %.2f cpu %.0f | local pin %.2f page %.2f | mate pin %.2f page %.2f (kernel %.2f)' % (
            f.split('bench_${TAG}_')[1][:-5], d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'], e['ms_per_step_median_rank0'], e['host_cpu_ms_per_step'],
            s['localalign_config2']['e2e_pinned_ms'], s['localalign_config2']['e2e_ms'], s['matealign_config4']['e2e_pinned_ms'], s['matealign_config4']['e2e_ms'], s['matealign_config4']['ms_per_step']))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
DFB_DEVICE_BUILD=1 DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_dev.txt; echo trace_rc=$?
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_host.txt; echo trace_rc=$?
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_simple.py 2> gpurun_out/trace_simple_${TAG}.txt; echo trace_rc=$?
grep -h "device:\|call 2" gpurun_out/trace_e2e_${TAG}_dev.txt gpurun_out/trace_e2e_${TAG}_host.txt gpurun_out/trace_simple_${TAG}.txt | grep -A12 "call 2 ----" | tail -60
