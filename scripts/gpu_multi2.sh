# 2-GPU box (gpurun --gpus 2): the sharded tool path on distinct ordinals, then the 2-rank bench
TAG=${1:-r03_n2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_tools_gpu.py -m gpu -x -q -k "distinct_gpus or sharded_over_contexts" --timeout 300 --timeout-method thread > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_rc=$?
tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$TAG.err
