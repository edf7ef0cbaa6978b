TAG=${1:-r04e}
mkdir -p gpurun_out
./gpurun_variants/poolgrow > gpurun_out/poolgrow_$TAG.txt 2>&1; cat gpurun_out/poolgrow_$TAG.txt
