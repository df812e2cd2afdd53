"""Kernel-level timings of the arena path at several batch sizes (CUDA events, L2 flushed when the
working set is small).  Usage: python scripts/kbench.py [N ...]   -> one JSON line per (N, S)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import ArenaConfig, BatchedBattleground  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, flush, iters=30):
    st = torch.cuda.current_stream()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3     # median, us


def run(N, S, bot, lcap, warm_frames):
    bg = BatchedBattleground(N, ships={bot: S}, config=ArenaConfig(laser_cap=lcap), seed=3)
    maps = torch.empty((N, 2, 5000), dtype=torch.int32, device=bg.device)
    for _ in range(warm_frames):
        bg.frame()
    from bench import L2Flush
    flush = L2Flush(bg.device, FLUSH_MODE)
    live = int(bg.state(("n_lasers",))["n_lasers"].sum().item())
    step_bytes = 88 * N * S + 64 * live
    t_bot = timeit(lambda: bg.request_actions(), flush)
    t_step = timeit(lambda: bg.generate_frame(), flush)
    t_step_hot = timeit(lambda: bg.generate_frame(), None)
    t_ras = timeit(lambda: bg.raster("bits", out=maps), flush)
    t_frame = timeit(lambda: (bg.frame(), bg.raster("bits", out=maps)), flush)
    t_fused = timeit(lambda: bg.frame(maps=maps), flush)
    out = {"flush": FLUSH_MODE, "N": N, "S": S, "bot": bot, "live_lasers_per_arena": live / N, "stride": bg.state_stride,
           "us": {"bots": t_bot, "step_cold": t_step, "step_hot_l2": t_step_hot, "raster": t_ras, "frame+raster": t_frame, "fused_frame": t_fused},
           "fused_env_steps_per_s": N / t_fused * 1e6, "fused_frac": (step_bytes + N * 40000) / t_fused / 1e3 / PEAK,
           "frame_knobs": {k: os.environ.get(k) for k in ("OFB_FRAME_TUNE", "OFB_FRAME_LPA", "OFB_FRAME_SW", "OFB_FRAME_NG", "OFB_FRAME_NBUF", "OFB_FRAME_K")},
           "step_alg_GBs": step_bytes / t_step / 1e3, "step_frac": step_bytes / t_step / 1e3 / PEAK,
           "raster_alg_GBs": N * 40000 / t_ras / 1e3, "raster_frac": N * 40000 / t_ras / 1e3 / PEAK,
           "env_steps_per_s": N / t_frame * 1e6}
    print(json.dumps(out), flush=True)
    del bg, maps, flush
    torch.cuda.empty_cache()


FLUSH_MODE = "write+read"

if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0].startswith("--flush="):
        FLUSH_MODE = args.pop(0).split("=", 1)[1]
    Ns = [int(a) for a in args] or [4096, 65536, 131072]
    for N in Ns:
        run(N, 7, "random", 0, 40)
    run(16384, 32, "stress", 2048, 12)
