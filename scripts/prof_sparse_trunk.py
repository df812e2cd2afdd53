import os, sys
sys.path.insert(0, "/root/repo")
import torch
from ofighters_b200 import BatchedBattleground
from ofighters_b200.policy import PolicyB200
N = 4096
bg = BatchedBattleground(N, ships={"random": 7}, seed=3)
maps = bg.raster("bits")
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    bg.frame(maps=maps)
pol = PolicyB200.random_init(seed=0, max_ships=N)
vec = bg.obs_vec[:, 0, :].contiguous()
for _ in range(3):
    pol.forward_argmax(maps, vec)
torch.cuda.synchronize()
print("done")
