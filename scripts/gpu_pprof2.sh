#!/bin/bash
mkdir -p gpurun_out
OFB_MAX_SHIPS=1024 python scripts/pbench.py 1024 tensor > gpurun_out/plain_p.log 2>&1 &&
OFB_MAX_SHIPS=1024 ncu --set full --clock-control none --import-source on -k regex:'k_tc_conv' -s 12 -c 5 -o gpurun_out/prof_policy2 \
    python scripts/pbench.py 1024 tensor > gpurun_out/ncu_pfull.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_pfull.log
