"""Fingerprint of the fused tail's outputs (dense pointer map + decisions) on a fixed scene: run under different OFB_* switches and compare."""
import hashlib
import sys
import torch
sys.path.insert(0, ".")
from oracle import policy_torch as po
from ofighters_b200 import BatchedBattleground
from ofighters_b200.policy import PolicyB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 333
w = po.init_weights(5, randomize_bn=True)
bg = BatchedBattleground(n, ships={"random": 7}, seed=11)
for _ in range(30):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec[:, 0, :].contiguous()
pol = PolicyB200(w, max_ships=max(16, n))
r = pol.forward(maps, vec, 1, want_ptr=True)
torch.cuda.synchronize()
print("ptr", hashlib.sha1(r["ptr"].cpu().numpy().tobytes()).hexdigest()[:16], "xy", hashlib.sha1(r["xy"].cpu().numpy().tobytes()).hexdigest()[:16],
      "nan", int(torch.isnan(r["ptr"]).sum()))
