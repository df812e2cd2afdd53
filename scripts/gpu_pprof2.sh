#!/bin/bash
# One ncu --set full capture of the forward's four product kernels on an 8192-ship chunk at the laser peak (after the same command
# has exited 0 without ncu).
O=gpurun_out/pf; mkdir -p $O
OFB_MAX_SHIPS=8192 python scripts/pbench.py 8192 tensor > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
tail -1 $O/plain.log
OFB_MAX_SHIPS=8192 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_st_trunk12|k_tz_tail|k_heads|k_tc_dense1' -s 12 -c 4 \
    -o $O/prof_policy python scripts/pbench.py 8192 tensor > $O/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/ncu_full.log
