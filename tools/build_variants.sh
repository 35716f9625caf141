#!/bin/bash
# build tuning variants of the CUDA library: tools/build_variants.sh name "-DFLAGS" ...
cd "$(dirname "$0")/../microclimate.jl_b200/csrc"
mkdir -p variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared -Xcompiler -fPIC $flags -o variants/lib_$name.so mc_api.cu &
done
wait
ls -la variants
