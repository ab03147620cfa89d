#!/bin/bash
# Builds liblumo_gpu.so with extra -D flags into lumo_b200/variants/ (A/B timing on the GPU box: LUMO_GPU_SO=<path> selects it).
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"
set -e
cd "$(dirname "$0")/.."
mkdir -p lumo_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -std=c++17 -Xcompiler -fPIC -shared -I include $2 \
  -o lumo_b200/variants/liblumo_gpu_$1.so lumo_b200/csrc/gpu/lumo_gpu.cu
echo built $1
