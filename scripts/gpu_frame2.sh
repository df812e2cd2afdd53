#!/bin/bash
mkdir -p gpurun_out/fr
timeout 900 python -m pytest tests/test_gpu_arena.py -m gpu -x -q -k "fused or host_tape" > gpurun_out/fr/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/fr/pytest.log
for c in "16 4 2 2" "8 4 2 2" "32 4 2 2" "16 4 3 1" "8 4 3 1"; do
  set -- $c
  OFB_FRAME_LPA=$1 OFB_FRAME_SW=$2 OFB_FRAME_NG=$3 OFB_FRAME_NBUF=$4 timeout 300 python scripts/kbench.py 4096 131072 2>&1 | head -2 | grep -o '"N": [0-9]*\|fused_frame": [0-9.]*\|fused_frac": [0-9.]*' | tr "\n" " "; echo " <- $c"
done | tee gpurun_out/fr/kbench2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 1 -o gpurun_out/fr/prof_frame131k_b python scripts/prof_frame.py 131072 > gpurun_out/fr/ncu131k.log 2>&1
tail -2 gpurun_out/fr/ncu131k.log
