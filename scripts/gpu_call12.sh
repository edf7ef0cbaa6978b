# parity suite + bench of the current build; A/B: probe kernel at six CTAs per SM
TAG=${1:-r03o}
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_long_windows.py -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -4 gpurun_out/pytest_$TAG.log
QUICK="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary"
for rep in 1 2; do
timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_base_$rep.json 2> gpurun_out/bench_${TAG}_base_$rep.err; echo base_rc=$?
DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_p6.so timeout 300 python bench.py $QUICK > gpurun_out/bench_${TAG}_p6_$rep.json 2> gpurun_out/bench_${TAG}_p6_$rep.err; echo p6_rc=$?
done
