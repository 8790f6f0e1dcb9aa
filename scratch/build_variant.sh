#!/bin/bash
# usage: scratch/build_variant.sh NAME [-DPART_NT=128 ...]   -> scratch/variants/lib_NAME.so (config-2 kernels only)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
/usr/local/cuda/bin/nvcc -std=c++17 -O3 --fmad=false -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared -DRIMU_TUNE_ONLY_MOM1D "$@" rimu.jl_b200/csrc/api.cu rimu.jl_b200/csrc/sort.cu \
  -o scratch/variants/lib_$name.so -ldl
echo built scratch/variants/lib_$name.so
