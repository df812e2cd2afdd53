"""Policy forward timings (CUDA events): python scripts/pbench.py [n_arenas] [engine ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import BatchedBattleground  # noqa: E402
from ofighters_b200.policy import PolicyB200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    engines = sys.argv[2:] or ["tensor", "cuda_core"]
    bg = BatchedBattleground(n, ships={"random": 7}, seed=5)
    for _ in range(30):
        bg.frame()
    maps = bg.raster("bits")
    vec = bg.obs_vec[:, 0, :].contiguous()
    pol = PolicyB200.random_init(device=bg.device, seed=0, max_ships=int(os.environ.get('OFB_MAX_SHIPS', '4096')))
    if os.environ.get("OFB_TZ_DEBUG"):
        import ctypes
        from ofighters_b200 import _lib
        dbg = torch.zeros(16 + 8 * 14, dtype=torch.int64, device=bg.device)
        lib = _lib.load()
        lib.ofb_policy_tz_debug.argtypes = [ctypes.c_void_p]
        lib.ofb_policy_tz_debug3.argtypes = [ctypes.c_void_p]
        if os.environ.get("OFB_TZ_DEBUG") == "up3":
            lib.ofb_policy_tz_debug3(ctypes.c_void_p(dbg.data_ptr()))
        else:
            lib.ofb_policy_tz_debug(ctypes.c_void_p(dbg.data_ptr()))
        pol.set_engine("tensor")
        pol.forward_argmax(maps, vec)
        pol.forward_argmax(maps, vec)
        torch.cuda.synchronize()
        if os.environ.get("OFB_TZ_DEBUG") == "trunk":
            dbg.zero_()
            pol.forward(maps[:8192], vec[:8192], want_act=False, want_argmax=False)     # trunk + heads only
            torch.cuda.synchronize()
        st = dbg.cpu().tolist()
        print("tz_up4 stamps (cycles since start):", [x - st[0] if x else None for x in st[:13]])
        for k in range(14):
            row = st[16 + 8 * k: 16 + 8 * k + 7]
            base = st[0] or st[16]
            print("  work %2d: j0 %s j1 %s j2 %s j3 %s j4 %s j5 %s j6 %s" % ((k,) + tuple(x - base if x else None for x in row)))
        lib.ofb_policy_tz_debug(None)
        lib.ofb_policy_tz_debug3(None)
    for eng in engines:
        pol.set_engine(eng)
        for _ in range(2):
            pol.forward_argmax(maps, vec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 3
        a.record()
        for _ in range(iters):
            pol.forward_argmax(maps, vec)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        pol.profile(True)
        pol.forward_argmax(maps, vec)
        prof = pol.profile(False)
        print(json.dumps({"engine": eng, "layers_ms": {k: round(v, 3) for k, v in prof.items()}, "n_arenas": n, "ms_per_forward_batch": ms, "forwards_per_s": n / ms * 1e3,
                          "dense_equiv_TFLOPs": n * 155.3e6 / (ms * 1e-3) / 1e12}), flush=True)


if __name__ == "__main__":
    main()
