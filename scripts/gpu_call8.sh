# launch list (durations) of the end-to-end call, plus the fused tool and pipelined tests
TAG=${1:-r03j}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tools_gpu.py tests/test_tools_downstream.py -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest_$TAG.log
timeout 200 python scripts/gpu_trace_e2e.py > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_e2e_$TAG.csv python scripts/gpu_trace_e2e.py > gpurun_out/ncu_launch_$TAG.log 2>&1
echo ncu_launch_rc=$?
