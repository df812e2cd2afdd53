#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (default + policy workload), reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
python bench.py --steps 400 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --workload policy --steps 10 --warmup 3 > gpurun_out/bench_policy.json 2> gpurun_out/bench_policy.err; echo "bench policy rc=$?"
tail -3 gpurun_out/bench_policy.err; cat gpurun_out/bench_policy.json
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
