#!/bin/bash
O=gpurun_out/tr; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log
timeout 900 python bench.py --workload sharded1m --steps 60 --warmup 3 > $O/bench_sharded1m.json 2> $O/bench_sharded1m.err; echo "sharded1m rc=$?"; tail -3 $O/bench_sharded1m.err; cut -c1-1500 $O/bench_sharded1m.json
timeout 900 python bench.py --workload policy7 --steps 10 --warmup 3 > $O/bench_policy7.json 2> $O/bench_policy7.err; echo "policy7 rc=$?"; tail -3 $O/bench_policy7.err; cut -c1-700 $O/bench_policy7.json
timeout 900 python bench.py --workload policy --steps 10 --warmup 3 > $O/bench_policy.json 2> $O/bench_policy.err; echo "policy rc=$?"; tail -3 $O/bench_policy.err; cut -c1-500 $O/bench_policy.json
