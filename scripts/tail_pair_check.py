"""CTA-pair tail (cta_group::2) against the single-CTA tail: dense pointer maps and decisions must be identical."""
import os
import sys
import torch
sys.path.insert(0, ".")
from oracle import policy_torch as po
from ofighters_b200 import BatchedBattleground
from ofighters_b200.policy import PolicyB200

w = po.init_weights(5, randomize_bn=True)
for n in (1, 2, 5, 148, 149, 333):
    bg = BatchedBattleground(n, ships={"random": 7}, seed=11 + n)
    for _ in range(30):
        bg.frame()
    maps = bg.raster("bits")
    vec = bg.obs_vec[:, 0, :].contiguous()
    os.environ["OFB_POLICY_TAIL_PAIR"] = "0"
    single = PolicyB200(w, max_ships=max(16, n))
    os.environ["OFB_POLICY_TAIL_PAIR"] = "1"
    pair = PolicyB200(w, max_ships=max(16, n))
    rs = single.forward(maps, vec, 1, want_ptr=True)
    rp = pair.forward(maps, vec, 1, want_ptr=True)
    torch.cuda.synchronize()
    i1, xy1 = single.forward_argmax(maps, vec, 1)
    i2, xy2 = pair.forward_argmax(maps, vec, 1)
    torch.cuda.synchronize()
    print("n = %3d: ptr equal %s (max|diff| %.3e), xy equal %s, argmax-only xy equal %s" % (
        n, bool(torch.equal(rs["ptr"], rp["ptr"])), float((rs["ptr"] - rp["ptr"]).abs().max()), bool(torch.equal(rs["xy"], rp["xy"])),
        bool(torch.equal(xy1, xy2))))
