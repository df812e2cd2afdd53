#!/bin/bash
mkdir -p gpurun_out/tr
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/tr/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tr/pytest.log
python scripts/tbench.py --cpu 2>&1 | tail -1 | tee gpurun_out/tr/tbench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_tr_ -s 1200 -c 400 --csv --log-file gpurun_out/tr/launches.csv python scripts/tbench.py > gpurun_out/tr/ncu.log 2>&1; echo "ncu rc=$?"
