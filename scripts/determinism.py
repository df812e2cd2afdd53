"""Race hunt without a sanitizer: the same forward 12 times on the same inputs (default arenas at the laser peak + stress + empty
arenas, so that CTAs alternate between the trunk's compact path and its fallback) must give bit-identical flat / act / xy."""
import sys
import torch
sys.path.insert(0, ".")
from ofighters_b200 import ArenaConfig, BatchedBattleground
from ofighters_b200.policy import PolicyB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
bg = BatchedBattleground(n, ships={"random": 7}, seed=5)
for _ in range(30):
    bg.frame()
st = BatchedBattleground(40, ships={"stress": 32}, config=ArenaConfig(laser_cap=2048), seed=12)
for _ in range(12):
    st.frame()
m0, ms = bg.raster("bits"), st.raster("bits")
idx = torch.randperm(n + 40 + 20, generator=torch.Generator().manual_seed(1))
maps = torch.cat([m0, ms, torch.zeros((20, 2, 5000), dtype=torch.int32, device=bg.device)])[idx.to(bg.device)].contiguous()
vec = torch.cat([bg.obs_vec[:, 0, :], st.obs_vec[:, 0, :], bg.obs_vec[:20, 0, :]])[idx.to(bg.device)].contiguous()
for pair in (False, True):
    pol = PolicyB200.random_init(device=bg.device, seed=0, max_ships=maps.shape[0], tail_pair=pair)
    ref = None
    bad = 0
    for it in range(12):
        r = pol.forward(maps, vec, 1)
        flat = pol.debug_tap(3, maps.shape[0], (25, 25, 8)).clone()
        cur = (flat, r["act"].clone(), r["xy"].clone(), r["iaction"].clone())
        torch.cuda.synchronize()
        if ref is None:
            ref = cur
        elif not all(torch.equal(a, b) for a, b in zip(ref, cur)):
            bad += 1
    print("tail_pair=%s: %d arenas, 12 forwards, %d differ from the first" % (pair, maps.shape[0], bad))
