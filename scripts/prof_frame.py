"""Tiny driver for `ncu --set full`: a few fused-frame launches on N arenas at steady state (no timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import ArenaConfig, BatchedBattleground  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 7
bot = sys.argv[3] if len(sys.argv) > 3 else "random"
bg = BatchedBattleground(N, ships={bot: S}, config=ArenaConfig(laser_cap=2048 if S == 32 else 0), seed=3)
maps = torch.empty((N, 2, 5000), dtype=torch.int32, device=bg.device)
for _ in range(40 if S != 32 else 12):
    bg.frame(maps=maps)
torch.cuda.synchronize()
print("done", N, S, bot)
