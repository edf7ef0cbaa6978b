# fold-early pair loop (DFB_FOLD_EARLY=1, the in-tree build) against the previous fold order (gpurun_variants/..._fe0.so):
# full kernel parity with the new build, then an A/B of the resident step and the end-to-end call
TAG=${1:-r04b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded"
for rep in 1 2; do
  timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_fe1_$rep.json 2> gpurun_out/bench_${TAG}_fe1_$rep.err; echo fe1_rc=$?
  DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_fe0.so timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_fe0_$rep.json 2> gpurun_out/bench_${TAG}_fe0_$rep.err; echo fe0_rc=$?
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']; r=d['roofline']
        print(f.split('bench_')[1], 'value %.0f ms %.2f | sweep %.2f probe %.2f frac %.3f own %.3f | e2e %.0f ms %.2f min %.2f | h2d %.0f MB d2h %.0f MB | pool high %.2f GB' % (
            d['value'], d['ms_per_step'], r['kernel_ms'], r['probe_sweep_ms'], r['frac'], r['frac_own_minimum'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'],
            e['h2d_bytes_per_step']/1e6, e['d2h_bytes_per_step']/1e6, e['device_pool_used_high_bytes']/1e9))
        for k,v in d.get('secondary',{}).items(): print('   ',k,{a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in('workload',)})
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
