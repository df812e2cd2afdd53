"""The forward as one chunk must equal the forward in chunks (P = 7 policy ships per arena, 16384 arenas)."""
import sys
import torch
sys.path.insert(0, ".")
from ofighters_b200 import BatchedBattleground
from ofighters_b200.policy import PolicyB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bg = BatchedBattleground(n, ships={"random": 7}, seed=5)
for _ in range(30):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec.reshape(-1, 8).contiguous()
res = {}
for ms in (8192, n * 7):
    pol = PolicyB200.random_init(device=bg.device, seed=0, max_ships=ms)
    i, xy = pol.forward_argmax(maps, vec, 7)
    torch.cuda.synchronize()
    res[ms] = (i.clone(), xy.clone())
    print("max_ships", ms, "iaction sum", int(i.sum()), "xy sum", int(xy.sum()))
    del pol
a, b = res[8192], res[n * 7]
print("iaction equal:", bool(torch.equal(a[0], b[0])), "xy equal:", bool(torch.equal(a[1], b[1])))
if not torch.equal(a[1], b[1]):
    bad = (a[1] != b[1]).any(dim=1).nonzero().flatten()
    print("first differing ships:", bad[:20].tolist(), "count", int(bad.numel()))
