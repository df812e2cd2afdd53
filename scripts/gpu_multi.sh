#!/bin/bash
O=gpurun_out/m2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 400 --warmup 10 > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?"; tail -2 $O/bench2.err; cut -c1-400 $O/bench2.json
timeout 900 $TR bench.py --gpus 2 --workload sharded1m --steps 60 --warmup 3 > $O/sharded2.json 2> $O/sharded2.err; echo "sharded2 rc=$?"; tail -2 $O/sharded2.err; cut -c1-1300 $O/sharded2.json
timeout 900 $TR bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $O/ref2.json 2> $O/ref2.err; echo "ref2 rc=$?"; cut -c1-200 $O/ref2.json
