#!/bin/bash
# ncu evidence for the arena path: launch list of the bench command + full capture of step and raster.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python scripts/kbench.py > gpurun_out/kbench.jsonl 2> gpurun_out/kbench.err; cat gpurun_out/kbench.jsonl; tail -3 gpurun_out/kbench.err
python bench.py --steps 20 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_raster' -s 20 -c 4 -o gpurun_out/prof_arena \
    python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
