#!/bin/bash
# tools/build_variant.sh <name> <nvcc -D flags...>  ->  build_variants/librtb_cuda_<name>.so   (experiment builds; select with RTB_CUDA_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static "$@" \
  rendering_b200/csrc/cuda/rtb_api.cu -o build_variants/librtb_cuda_$name.so
