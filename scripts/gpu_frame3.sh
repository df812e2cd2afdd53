#!/bin/bash
mkdir -p gpurun_out/fr
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 1 -o gpurun_out/fr/prof_frame131k_c python scripts/prof_frame.py 131072 > gpurun_out/fr/ncu131k.log 2>&1
tail -2 gpurun_out/fr/ncu131k.log
