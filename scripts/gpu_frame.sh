#!/bin/bash
# Fused frame kernel: parity tests, then kernel timings for the tuning knobs.
mkdir -p gpurun_out/fr
timeout 600 python -m pytest tests/test_gpu_arena.py -m gpu -x -q -k "fused or host_tape" > gpurun_out/fr/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/fr/pytest.log
for cfg in "16 4 2 2" "16 4 1 3" "16 4 3 1" "16 8 2 2" "8 4 2 2" "32 4 2 2" "32 8 2 2"; do
  set -- $cfg
  OFB_FRAME_LPA=$1 OFB_FRAME_SW=$2 OFB_FRAME_NG=$3 OFB_FRAME_NBUF=$4 timeout 300 python scripts/kbench.py 4096 131072 2>&1 | head -2 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('$cfg', d['N'], {k: round(v, 1) for k, v in d['us'].items()}, 'fused_frac', round(d['fused_frac'], 3))
"
done | tee gpurun_out/fr/kbench.txt
