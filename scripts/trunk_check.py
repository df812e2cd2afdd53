"""Sparse tensor trunk (k_st_trunk12) against the CUDA-core sparse trunk and the dense tcgen05 trunk: pool2 taps."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import policy_torch as po
from ofighters_b200 import ArenaConfig, BatchedBattleground
from ofighters_b200.policy import PolicyB200

w = po.init_weights(5, randomize_bn=True)
scenes = []
bg = BatchedBattleground(600, ships={"random": 7}, seed=11)
scenes.append(("frame0", bg.raster("bits").clone(), bg.obs_vec[:, 0, :].clone()))
for _ in range(30):
    bg.frame()
scenes.append(("frame30", bg.raster("bits").clone(), bg.obs_vec[:, 0, :].clone()))
for _ in range(150):
    bg.frame()
scenes.append(("frame180", bg.raster("bits").clone(), bg.obs_vec[:, 0, :].clone()))
st = BatchedBattleground(300, ships={"stress": 32}, config=ArenaConfig(laser_cap=2048), seed=12)
for _ in range(12):
    st.frame()
scenes.append(("stress12", st.raster("bits").clone(), st.obs_vec[:, 0, :].clone()))
scenes.append(("empty", torch.zeros_like(scenes[0][1][:5]), scenes[0][2][:5].clone()))
full = torch.full_like(scenes[0][1][:3], -1)
scenes.append(("all-ones", full, scenes[0][2][:3].clone()))
pols = {"fused": PolicyB200(w, max_ships=600), "st12": PolicyB200(w, max_ships=600, fused_trunk=False),
        "cc": PolicyB200(w, max_ships=600, cc_sparse_trunk=True), "dense": PolicyB200(w, max_ships=600, dense_trunk=True)}
pols["fused"].set_taps(True)
for name, maps, vec in scenes:
    n = maps.shape[0]
    taps = {}
    for k, p in pols.items():
        p.forward(maps.contiguous(), vec.contiguous(), 1, want_act=False, want_argmax=False)
        torch.cuda.synchronize()
        taps[k] = [p.debug_tap(1, n, (100, 100, 8)).float().cpu(), p.debug_tap(2, n, (50, 50, 8)).float().cpu(),
                   p.debug_tap(3, n, (25, 25, 8)).float().cpu()]
    for k in ("fused", "st12", "dense"):
        msg = []
        for lvl, nm in enumerate(("pool2", "pool3", "pool4")):
            d = (taps[k][lvl] - taps["cc"][lvl]).abs()
            msg.append("%s max|diff| %.2e (%d values)" % (nm, float(d.max()), int((d > 0).sum())))
        print("%-9s %-5s vs cc: %s" % (name, k, "; ".join(msg)))
