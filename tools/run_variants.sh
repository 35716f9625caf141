#!/bin/bash
# run tools/gpu_profile.py once per variant library, per MC_SMEM_PAD in $PADS and per MC_BLOCKS_PER_SM in $BPSM
cd "$(dirname "$0")/.."
for f in microclimate.jl_b200/csrc/variants/lib_*.so; do
  for pad in ${PADS:-0}; do
    for b in ${BPSM:-0}; do
      echo "== $f pad=$pad blocks_per_sm=$b"; MC_BLOCKS_PER_SM=$b MC_SMEM_PAD=$pad MC_LIB=$PWD/$f timeout 300 python tools/gpu_profile.py 2>&1 | tail -1
    done
  done
done
