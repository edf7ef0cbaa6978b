# pageable uploads: workers per staging block (DFB_STAGE_SHIFT 19 = a thread per 512 KB, 20 = per MB, 18 = per 256 KB)
TAG=${1:-r04l}
mkdir -p gpurun_out
QUICK="--steps 5 --warmup 3 --no-cpu-baseline --no-sharded"
for rep in 1 2; do
for sh in 20 19 18; do
  DFB_STAGE_SHIFT=$sh timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_s${sh}_$rep.json 2> gpurun_out/bench_${TAG}_s${sh}_$rep.err; echo s${sh}_rc=$?
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']; s=d.get('secondary',{})
        print('%-10s e2e %.2f | local pin %.2f page %.2f | mate pin %.2f page %.2f' % (
            f.split('bench_${TAG}_')[1][:-5], e['ms_per_step'],
            s['localalign_config2']['e2e_pinned_ms'], s['localalign_config2']['e2e_ms'], s['matealign_config4']['e2e_pinned_ms'], s['matealign_config4']['e2e_ms']))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
