# builds gpurun_variants/libdefuse_b200_<name>.so with extra compiler flags, for A/B runs inside one gpurun call
# (select it with DFB_LIB_PATH).  usage: bash scripts/build_variant.sh <name> <flags...>   e.g.  fe0 -DDFB_FOLD_EARLY=0
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-fvisibility=hidden "$@" \
     -c defuse_b200/csrc/dfb_api.cu -o gpurun_variants/dfb_api_$NAME.o
g++ -shared -o gpurun_variants/libdefuse_b200_$NAME.so gpurun_variants/dfb_api_$NAME.o defuse_b200/build/dfb_micro.o \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
rm -f gpurun_variants/dfb_api_$NAME.o
ls -la gpurun_variants/libdefuse_b200_$NAME.so
