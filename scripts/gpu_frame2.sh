#!/bin/bash
mkdir -p gpurun_out/fr
for c in "16 4 2 2" "16 8 2 1" "16 12 2 1" "8 8 2 1" "32 12 2 1"; do
  set -- $c
  OFB_FRAME_LPA=$1 OFB_FRAME_SW=$2 OFB_FRAME_NG=$3 OFB_FRAME_NBUF=$4 timeout 300 python scripts/kbench.py 4096 16384 131072 2>&1 | head -3 | grep -o '"N": [0-9]*\|fused_frame": [0-9.]*\|fused_frac": [0-9.]*' | tr "\n" " "; echo " <- $c"
done | tee gpurun_out/fr/kbench7.txt
OFB_FRAME_LPA=16 OFB_FRAME_SW=12 OFB_FRAME_NG=2 OFB_FRAME_NBUF=1 python scripts/frame_phases.py 4096 | tail -3
