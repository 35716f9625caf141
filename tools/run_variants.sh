#!/bin/bash
# run tools/gpu_profile.py once per variant library (and per MC_SMEM_PAD in $PADS)
cd "$(dirname "$0")/.."
for f in microclimate.jl_b200/csrc/variants/lib_*.so; do
  for pad in ${PADS:-0}; do
    echo "== $f pad=$pad"; MC_SMEM_PAD=$pad MC_LIB=$PWD/$f timeout 300 python tools/gpu_profile.py 2>&1 | tail -1
  done
done
