# A/B of the end-to-end path on ONE box (host-dependent numbers differ from box to box): device-built vs host-built
# job lists, chunk counts; then the tools at scale with their phase traces
TAG=${1:-ab}
mkdir -p gpurun_out
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-secondary"
for rep in 1 2; do
for v in "dev8:" "host8:DFB_HOST_BUILD=1" "dev6:DFB_PIPELINE_CHUNKS=6" "host6:DFB_HOST_BUILD=1 DFB_PIPELINE_CHUNKS=6" "dev5:DFB_PIPELINE_CHUNKS=5"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 200 python bench.py $Q > gpurun_out/ab_${TAG}_${name}_$rep.json 2> gpurun_out/ab_${TAG}_${name}_$rep.err; echo ${name}_$rep rc=$?
done
done
timeout 300 python scripts/gpu_tool_scale.py 4000 100 > gpurun_out/tool_scale_split_$TAG.json 2> gpurun_out/tool_scale_split_$TAG.err; echo scale_split_rc=$?
