# usage: bash scripts/gpu_ncu_full.sh <tag>   -- the one ncu session of a gpurun call: ncu --set full of the two sweeps on the
# small bench config (one 200 k-task launch per kernel), after the same command has exited 0 without ncu
TAG=${1:-x}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --clusters 2000 --no-cpu-baseline --no-secondary --no-sharded --e2e-steps 1"
timeout 200 $SMALL > gpurun_out/plain_full_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_(fast|probe)_kernel' -s 2 -c 2 -o gpurun_out/prof_$TAG $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo ncu_full_rc=$?
tail -3 gpurun_out/ncu_full_$TAG.log
