#!/bin/bash
# ncu --set full capture of the trunk kernel alone on an 8192-arena chunk at the laser peak (after a plain run has exited 0).
O=gpurun_out/pt; mkdir -p $O
OFB_MAX_SHIPS=8192 python scripts/pbench.py 8192 tensor > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
tail -1 $O/plain.log
OFB_MAX_SHIPS=8192 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_st_trunk12' -s 3 -c 1 \
    -o $O/prof_trunk python scripts/pbench.py 8192 tensor > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
