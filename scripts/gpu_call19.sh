# sweep-state arena: parity of the pipelined paths, bench (host- and device-built lists), DFB_TRACE timelines
TAG=${1:-r04g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 180 --timeout-method thread \
  -k "pipelined or overflow or full_size or large or staged or stress" > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary"
for rep in 1 2; do
timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_host_$rep.json 2> gpurun_out/bench_${TAG}_host_$rep.err; echo rc=$?
DFB_DEVICE_BUILD=1 timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_dev_$rep.json 2> gpurun_out/bench_${TAG}_dev_$rep.err; echo rc=$?
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']
        print('%-10s value %.0f ms %.2f | e2e %.0f mean %.2f min %.2f med %.2f cpu %.0f | h2d %.0f d2h %.0f MB | pool high %.2f GB' % (
            f.split('bench_${TAG}_')[1][:-5], d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'], e['ms_per_step_median_rank0'], e['host_cpu_ms_per_step'],
            e['h2d_bytes_per_step']/1e6, e['d2h_bytes_per_step']/1e6, e['device_pool_used_high_bytes']/1e9))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
DFB_DEVICE_BUILD=1 DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_dev.txt; echo trace_rc=$?
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_host.txt; echo trace_rc=$?
grep -h "device:\|call .*ms" gpurun_out/trace_e2e_${TAG}_dev.txt gpurun_out/trace_e2e_${TAG}_host.txt | tail -44
