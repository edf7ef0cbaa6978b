# usage: bash scripts/gpu_ncu_launch.sh <tag>   -- the one ncu session of a gpurun call: the launch list (durations) of the
# small bench config, after the same command has exited 0 without ncu
TAG=${1:-x}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --clusters 2000 --no-cpu-baseline --no-secondary --no-sharded --e2e-steps 1"
timeout 200 $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
echo ncu_launch_rc=$?
