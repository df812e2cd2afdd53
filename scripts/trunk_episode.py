"""trunk12 time of the dense (tcgen05) and the sparse (CUDA-core, default) kernels along one 200-frame episode."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import BatchedBattleground  # noqa: E402
from ofighters_b200.policy import PolicyB200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bg = BatchedBattleground(N, ships={"random": 7}, seed=3)
maps = bg.raster("bits")
os.environ["OFB_POLICY_DENSE_TRUNK"] = "1"
dense = PolicyB200.random_init(seed=0, max_ships=N)
os.environ.pop("OFB_POLICY_DENSE_TRUNK", None)
sparse = PolicyB200.random_init(seed=0, max_ships=N)
rows = []
for t in range(200):
    if t % 20 == 0:
        vec = bg.obs_vec[:, 0, :].contiguous()
        r = {"t": t, "lasers": float(bg.state(("n_lasers",))["n_lasers"].float().mean()),
             "alive": float(bg.state(("ship_alive",))["ship_alive"].float().sum(dim=1).mean())}
        for name, pol in (("dense", dense), ("sparse", sparse)):
            pol.forward_argmax(maps, vec)
            pol.profile(True)
            for _ in range(3):
                pol.forward_argmax(maps, vec)
            r[name] = pol.profile(False)["trunk12"]
        rows.append(r)
        print(json.dumps(r), flush=True)
    bg.frame(maps=maps)
print(json.dumps({"mean_dense_ms": sum(r["dense"] for r in rows) / len(rows), "mean_sparse_ms": sum(r["sparse"] for r in rows) / len(rows), "arenas": N}))
