#!/bin/bash
mkdir -p gpurun_out/tr
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/tr/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/tr/pytest.log
