# usage: bash scripts/gpu_bench_profile.sh <tag>     (runs under gpurun; one GPU)
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 1 --clusters 2000 --no-cpu-baseline --no-secondary --e2e-steps 1"
$SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
echo ncu_launch_rc=$?
$SMALL > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dp_fast_kernel -s 2 -c 2 -o gpurun_out/prof_$TAG $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo ncu_full_rc=$?
tail -3 gpurun_out/ncu_full_$TAG.log
