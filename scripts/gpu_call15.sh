# DFB_TRACE timelines (host laps + device events per chunk): split one-call path with host- and device-built job lists,
# simple one-call path at the matealign and localalign shapes
TAG=${1:-r04c}
mkdir -p gpurun_out
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_host.txt; echo trace_rc=$?
DFB_DEVICE_BUILD=1 DFB_TRACE=1 timeout 200 python scripts/gpu_trace_e2e.py 2> gpurun_out/trace_e2e_${TAG}_dev.txt; echo trace_rc=$?
DFB_TRACE=1 timeout 200 python scripts/gpu_trace_simple.py 2> gpurun_out/trace_simple_${TAG}.txt; echo trace_rc=$?
grep -h "device:\|call 2" gpurun_out/trace_e2e_${TAG}_dev.txt gpurun_out/trace_simple_${TAG}.txt | tail -40
