# usage: bash scripts/gpu_ncu.sh <tag>   -- ncu --set full of the DP kernels on the small bench config
TAG=${1:-x}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --clusters 2000 --no-cpu-baseline --no-secondary --e2e-steps 1"
timeout 200 $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dp_fast_kernel -s 2 -c 2 -o gpurun_out/prof_$TAG $SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
echo ncu_full_rc=$?
tail -3 gpurun_out/ncu_full_$TAG.log
