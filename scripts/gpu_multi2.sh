#!/bin/bash
# Multi-GPU pass, the driver's commands (one process per GPU, NCCL): default bench (= sharded1m with learning at N > 1) and the reference arm.
N=${1:-2}
O=gpurun_out/mm$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err; cut -c1-330 $O/bench.json
timeout 900 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/ref.json 2> $O/ref.err; echo "ref rc=$?"; cut -c1-200 $O/ref.json
