# bounds-check build of HEAD under the GPU parity suite (compute-sanitizer is closed on the pool), then the chunk count of
# the split one-call path once more now that the head of a batch is shorter
TAG=${1:-r04k}
mkdir -p gpurun_out
DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_check.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 240 --timeout-method thread > gpurun_out/pytest_check_$TAG.log 2>&1; echo pytest_check_rc=$?
tail -3 gpurun_out/pytest_check_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-secondary"
for K in 4 5 6 7 8; do
  DFB_PIPELINE_CHUNKS=$K timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_K$K.json 2> gpurun_out/bench_${TAG}_K$K.err; echo K${K}_rc=$?
done
timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo default_rc=$?
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']; s=d.get('sharded_merge') or {}
        print('%-10s value %.0f ms %.2f | e2e %.0f mean %.2f min %.2f med %.2f cpu %.0f | pool high %.2f GB | sharded merge %.1f ms align %.1f' % (
            f.split('bench_${TAG}_')[1][:-5], d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['ms_per_step_min_rank0'], e['ms_per_step_median_rank0'], e['host_cpu_ms_per_step'],
            e['device_pool_used_high_bytes']/1e9, s.get('merge_ms',-1), s.get('align_ms_max_over_ranks',-1)))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
