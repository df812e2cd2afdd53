#!/bin/bash
# Multi-GPU pass (one process per GPU, NCCL): default bench, sharded1m with learning, reference arm.  Usage: gpu_multi.sh N
N=${1:-2}
O=gpurun_out/m$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 400 --warmup 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err; cut -c1-330 $O/bench.json
timeout 900 $TR bench.py --gpus $N --workload sharded1m --steps 60 --warmup 3 > $O/sharded.json 2> $O/sharded.err; echo "sharded rc=$?"; tail -2 $O/sharded.err; cut -c1-330 $O/sharded.json
timeout 900 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 3 > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"; cut -c1-200 $O/ref.json
