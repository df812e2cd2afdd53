#!/bin/bash
# Full single-GPU pass: parity tests, smoke, the default bench (= the metric's config) and the reference arm, then ONE profiler
# pass: the ncu launch list of the same bench command (only after it has exited 0 without ncu).
O=gpurun_out/check; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; rc=$?; echo "bench rc=$rc"; tail -3 $O/bench.err; cut -c1-400 $O/bench.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/bench_ref.json
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 5 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
