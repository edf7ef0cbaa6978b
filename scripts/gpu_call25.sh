# probe sweep with the next queue batch claimed and prefetched early (gpurun_variants/..._pf1.so, -DDFB_PROBE_PREFETCH=1)
# against the in-tree build: parity of both, then an A/B of the resident step
TAG=${1:-r04p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 240 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_main_rc=$?
tail -2 gpurun_out/pytest_$TAG.log
DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_pf1.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 240 --timeout-method thread \
  -k "split or pipelined or stress or large or random or long or overflow or full_size" > gpurun_out/pytest_${TAG}_pf1.log 2>&1; echo pytest_pf1_rc=$?
tail -2 gpurun_out/pytest_${TAG}_pf1.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary"
for rep in 1 2; do
  timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_main_$rep.json 2> gpurun_out/bench_${TAG}_main_$rep.err; echo main_rc=$?
  DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_pf1.so timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_pf1_$rep.json 2> gpurun_out/bench_${TAG}_pf1_$rep.err; echo pf1_rc=$?
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']; r=d['roofline']
        print('%-8s value %.0f ms %.3f | sweep %.3f probe %.3f | e2e %.2f min %.2f' % (f.split('bench_${TAG}_')[1][:-5], d['value'], d['ms_per_step'], r['kernel_ms'], r['probe_sweep_ms'], e['ms_per_step'], e['ms_per_step_min_rank0']))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
