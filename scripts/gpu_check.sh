#!/bin/bash
# Full GPU pass: parity tests, smoke, benches, ncu launch list + full capture of the fused frame kernel.
O=gpurun_out/check; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 900 python bench.py --steps 1000 --warmup 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err; cut -c1-300 $O/bench.json
timeout 900 python bench.py --workload stress --steps 200 --warmup 10 > $O/bench_stress.json 2> $O/bench_stress.err; echo "bench stress rc=$?"; cut -c1-200 $O/bench_stress.json
timeout 900 python bench.py --workload policy --steps 10 --warmup 3 > $O/bench_policy.json 2> $O/bench_policy.err; echo "bench policy rc=$?"; cut -c1-200 $O/bench_policy.json
timeout 900 python bench.py --impl reference --steps 50 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err; cut -c1-200 $O/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 1 -o $O/prof_frame4096 python scripts/prof_frame.py 4096 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 30 -c 1 -o $O/prof_frame131k python scripts/prof_frame.py 131072 > $O/ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
