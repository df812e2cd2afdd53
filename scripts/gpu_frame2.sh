#!/bin/bash
mkdir -p gpurun_out/fr
for c in "16 4 2 2" "16 4 2 0" "16 8 2 0" "16 12 2 0" "16 8 3 0" "8 8 2 0" "32 12 2 0"; do
  set -- $c
  OFB_FRAME_LPA=$1 OFB_FRAME_SW=$2 OFB_FRAME_NG=$3 OFB_FRAME_NBUF=$4 timeout 300 python scripts/kbench.py 4096 16384 131072 2>&1 | head -3 | grep -o '"N": [0-9]*\|fused_frame": [0-9.]*\|fused_frac": [0-9.]*' | tr "\n" " "; echo " <- $c"
done | tee gpurun_out/fr/kbench8.txt
OFB_FRAME_LPA=16 OFB_FRAME_SW=8 OFB_FRAME_NG=2 OFB_FRAME_NBUF=0 python scripts/frame_phases.py 4096 | tail -3
