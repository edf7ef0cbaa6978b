# chunk views of the window table + early release of the sweep state: parity of the pipelined paths, then an A/B of the
# end-to-end call against the previous build (gpurun_variants/libdefuse_b200_head.so), host- and device-built job lists
TAG=${1:-r04a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method thread \
  -k "pipelined or overflow or full_size or planted or stress or staged" > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary"
for rep in 1 2; do
for build in host dev; do
  if [ $build = dev ]; then export DFB_DEVICE_BUILD=1; else unset DFB_DEVICE_BUILD; fi
  timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_new_${build}_$rep.json 2> gpurun_out/bench_${TAG}_new_${build}_$rep.err; echo new_${build}_rc=$?
  DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_head.so timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_head_${build}_$rep.json 2> gpurun_out/bench_${TAG}_head_${build}_$rep.err; echo head_${build}_rc=$?
done
done
unset DFB_DEVICE_BUILD
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']
        print(f.split('bench_')[1], 'value %.0f ms %.2f | e2e %.0f ms %.2f min %.2f med %.2f | cpu %.0f | h2d %.0f MB | pool high %.2f GB reserved %.2f GB' % (
            d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'], e['ms_per_step_median_rank0'],
            e['host_cpu_ms_per_step'], e['h2d_bytes_per_step']/1e6, e['device_pool_used_high_bytes']/1e9, e['device_pool_reserved_bytes']/1e9))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_host.txt; echo trace_rc=$?
DFB_DEVICE_BUILD=1 DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_dev.txt; echo trace_rc=$?
