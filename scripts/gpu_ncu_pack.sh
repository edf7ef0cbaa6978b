TAG=${1:-x}
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --clusters 20000 --no-cpu-baseline --no-secondary --e2e-steps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 0 -c 2 -o gpurun_out/prof_pack_$TAG $SMALL > gpurun_out/ncu_pack_$TAG.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_pack_$TAG.log
