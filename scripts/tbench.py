"""Timing of the Q-learning update path: Trainer.replay's pieces on the GPU (CUDA events) and the torch-CPU oracle beside it.
Usage: python scripts/tbench.py [--cpu]   -> one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofighters_b200 import BatchedBattleground  # noqa: E402
from ofighters_b200.trainer import TrainerB200  # noqa: E402

B = 8
bg = BatchedBattleground(B, ships={"random": 7}, seed=11)
for _ in range(35):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec[:, 0, :].contiguous()
g = torch.Generator().manual_seed(1)
ta = (torch.randn((B, 2), generator=g) * 3).cuda()
tp = (torch.randn((B, 400, 400), generator=g) * 0.5).cuda()
tr = TrainerB200(learning_rate=1e-4, batch_size=B)
st = torch.cuda.current_stream()


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


out = {"batch": B}
out["fit_ms"] = timeit(lambda: tr.fit(maps, vec, ta, tp, sync_model=False))
out["forward_train_ms"] = timeit(lambda: tr.forward_train(maps, vec))
out["predict_inference_ms"] = timeit(lambda: tr.predict(maps, vec))
t0 = time.perf_counter()
for _ in range(5):
    tr.sync_model()
torch.cuda.synchronize()
out["sync_model_ms_wall"] = (time.perf_counter() - t0) / 5 * 1e3
for k in range(B):
    tr.remember((maps[k], vec[k]), k % 2, (10 * k, 20 * k), 1.0, (maps[(k + 1) % B], vec[(k + 1) % B]), False)
t0 = time.perf_counter()
for _ in range(5):
    tr.replay(B)
torch.cuda.synchronize()
out["replay_ms_wall"] = (time.perf_counter() - t0) / 5 * 1e3
out["fit_samples_per_s"] = B / (out["fit_ms"] * 1e-3)
# algorithmic FLOPs of one fit: forward 155.3 MFLOP per sample, backward ~2x
out["fit_tflops"] = 3 * 155.3e6 * B / (out["fit_ms"] * 1e-3) / 1e12
if "--cpu" in sys.argv:
    from oracle import policy_torch as po
    from oracle import policy_train_torch as pt
    from tests.test_gpu_train import _dense_image
    w = po.init_weights(0)
    img, vh = _dense_image(maps), vec.cpu()
    opt = pt.KerasAdam()
    pt.fit(w, opt, img, vh, ta.cpu(), tp.cpu())
    t0 = time.perf_counter()
    for _ in range(3):
        pt.fit(w, opt, img, vh, ta.cpu(), tp.cpu())
    out["cpu_oracle_fit_ms"] = (time.perf_counter() - t0) / 3 * 1e3
    out["cpu_threads"] = torch.get_num_threads()
print(json.dumps(out))
