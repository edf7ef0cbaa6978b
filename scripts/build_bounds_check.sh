# builds gpurun_variants/libdefuse_b200_check.so: the library with -DDFB_BOUNDS_CHECK (every indexed global-memory access of
# the kernels range-checked; a violation turns the plan's fetch / sync into DFB_ERR_STATE).  Run the GPU suite against it with
#   DFB_LIB_PATH=$PWD/gpurun_variants/libdefuse_b200_check.so python -m pytest tests -m gpu -q
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-fvisibility=hidden -DDFB_BOUNDS_CHECK \
     -c defuse_b200/csrc/dfb_api.cu -o gpurun_variants/dfb_api_check.o
g++ -shared -o gpurun_variants/libdefuse_b200_check.so gpurun_variants/dfb_api_check.o defuse_b200/build/dfb_micro.o \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
rm -f gpurun_variants/dfb_api_check.o
ls -la gpurun_variants/libdefuse_b200_check.so
