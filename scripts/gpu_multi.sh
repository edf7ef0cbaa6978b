# usage: bash scripts/gpu_multi.sh <N> <tag>   (gpurun --gpus N)
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo rc=$?
cat gpurun_out/bench_${TAG}_n$N.json | cut -c1-600; tail -5 gpurun_out/bench_${TAG}_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/benchref_${TAG}_n$N.json 2> gpurun_out/benchref_${TAG}_n$N.err; echo rc=$?
cat gpurun_out/benchref_${TAG}_n$N.json | cut -c1-300
