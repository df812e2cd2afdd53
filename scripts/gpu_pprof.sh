#!/bin/bash
mkdir -p gpurun_out
python scripts/pbench.py 4096 > gpurun_out/pbench.jsonl 2> gpurun_out/pbench.err; cat gpurun_out/pbench.jsonl; tail -3 gpurun_out/pbench.err
python scripts/pbench.py 1024 tensor > gpurun_out/plain_p.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_policy.csv \
    python scripts/pbench.py 1024 tensor > gpurun_out/ncu_p.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_tc_conv|k_heads|k_dense1' -s 9 -c 7 -o gpurun_out/prof_policy \
    python scripts/pbench.py 1024 tensor > gpurun_out/ncu_pfull.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_pfull.log
