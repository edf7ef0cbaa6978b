# first-batch cost: pool presized in one step (default) against growth on demand (DFB_POOL_PRESIZE=0)
TAG=${1:-r04f}
mkdir -p gpurun_out
for mode in 1 0; do
DFB_POOL_PRESIZE=$mode DFB_DEVICE_BUILD=1 DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_dev_presize$mode.txt; echo trace_rc=$?
DFB_POOL_PRESIZE=$mode DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_host_presize$mode.txt; echo trace_rc=$?
DFB_POOL_PRESIZE=$mode DFB_TRACE=1 timeout 200 python scripts/gpu_trace_simple.py 2> gpurun_out/trace_simple_${TAG}_presize$mode.txt; echo trace_rc=$?
done
grep -H "call .*ms\|presized" gpurun_out/trace_*_${TAG}_*.txt
