"""Cycle breakdown of the fused frame kernel's warp roles (debug counters of ofb_debug_frame_prof)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import BatchedBattleground, _lib  # noqa: E402
from bench import L2Flush  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bg = BatchedBattleground(N, ships={"random": 7}, seed=3)
maps = torch.empty((N, 2, 5000), dtype=torch.int32, device=bg.device)
for _ in range(40):
    bg.frame(maps=maps)
flush = L2Flush(bg.device)
prof = torch.zeros(148 * 32 * 8 + 8, dtype=torch.int64, device=bg.device)
lib = _lib.load()
lib.ofb_debug_frame_prof(C.c_void_p(prof.data_ptr()))
flush()
bg.frame(maps=maps)
torch.cuda.synchronize()
lib.ofb_debug_frame_prof(None)
sec = prof[148 * 32 * 8:].cpu().numpy()
p = prof[:148 * 32 * 8].reshape(148, 32, 8).cpu().numpy()
if sec[7] > 0:
    print("step_tile sections, cycles per unit (issue+bot | wait load0 | setup | laser loop | ship move | rewards+append | store):", [int(x / sec[7]) for x in sec[:7]], "units", int(sec[7]))
sw = int(os.environ.get("OFB_FRAME_SW", "4"))
ng = int(os.environ.get("OFB_FRAME_NG", "2"))
st = p[:, :sw, :2]
print("N", N, "stepper warps: total cycles mean %.0f max %d, waiting for a slot mean %.0f" % (st[..., 0].mean(), st[..., 0].max(), st[..., 1].mean()))
import numpy as np
tot, umax, nun = p[:, :sw, 0].reshape(-1), p[:, :sw, 2].reshape(-1), p[:, :sw, 3].reshape(-1)
print("  stepper total cycles percentiles 50/90/99/max:", [int(np.percentile(tot, q)) for q in (50, 90, 99, 100)],
      " longest unit 50/90/99/max:", [int(np.percentile(umax, q)) for q in (50, 90, 99, 100)],
      " units per warp min/max:", int(nun.min()), int(nun.max()), " mean unit cycles:", round(float((tot - p[:, :sw, 1].reshape(-1)).sum() / nun.sum())))
ra = p[:, sw:sw + ng, :7].astype(float)
names = ["total", "drain+bar1", "zero+bar2", "wait full", "compose", "fence+bar3", "arenas"]
print("raster groups (mean over CTAs x groups):", {n: round(ra[..., j].mean(), 0) for j, n in enumerate(names)})
per = ra[..., 1:6].sum(axis=(0, 1)) / ra[..., 6].sum()
print("per arena cycles:", dict(zip(names[1:6], per.round(0))), "sum", per.sum().round(0))
