#!/bin/bash
# ncu --set full capture of the fused tail alone (single-CTA and CTA-pair form) on an 8192-ship chunk.
O=gpurun_out/ptl; mkdir -p $O
for P in 0 1; do
  OFB_POLICY_TAIL_PAIR=$P OFB_MAX_SHIPS=8192 python scripts/pbench.py 8192 tensor > $O/plain$P.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain$P.log; exit 1; }
  tail -1 $O/plain$P.log | cut -c1-200
  OFB_POLICY_TAIL_PAIR=$P OFB_MAX_SHIPS=8192 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tz_tail' -s 3 -c 1 \
      -o $O/prof_tail$P python scripts/pbench.py 8192 tensor > $O/ncu_full$P.log 2>&1
  echo "ncu full rc=$?"; tail -1 $O/ncu_full$P.log
done
