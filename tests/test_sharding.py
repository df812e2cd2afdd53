"""Multi-GPU path on CPU: world-size-2 gloo run of the sharding logic, plus the proof that sharded
results equal the single-process result arena-for-arena (arenas are independent and the RNG is keyed by
global arena id) -- here with the C oracle standing in for the device, same keying as the CUDA bots."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["OFB_ROOT"])
from ofighters_b200 import sharding
from oracle.step_c import ArenasC
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["OFB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
N, S, T, seed = 37, 7, 60, 0xBEEF
lo, hi = sharding.shard_range(N, rank, world)
c0 = ArenasC(np.zeros((hi - lo, S, 2), np.int32))
c = ArenasC(c0.random_spawn(seed, 0, arena0=lo))
for t in range(T):
    c.step(c.bot_actions("random", seed, t, arena0=lo))
c.reset(c.random_spawn(seed, 1, arena0=lo))                  # episode end: stats accumulate locally
stats = torch.from_numpy(c.arr["stats"].copy())
local = stats.clone()
sharding.reduce_episode_stats(stats)                         # K7: the path's only collective
flat = torch.full((10,), float(rank + 3), dtype=torch.float32)     # the shared trainer's weights live on rank 0
sharding.broadcast_weights(flat, src=0)
ls = torch.tensor([1.5 * (rank + 1), rank + 1.0], dtype=torch.float64)
sharding.reduce_loss_stats(ls)
np.savez(os.environ["OFB_OUT"] + ".%d.npz" % rank, lo=lo, hi=hi, local=local.numpy(), total=stats.numpy(),
         x=c.arr["ship_x"], y=c.arr["ship_y"], score=c.arr["ship_reward"], flat=flat.numpy(), ls=ls.numpy())
dist.destroy_process_group()
'''


def test_shard_ranges_tile_the_arena_set():
    from ofighters_b200.sharding import shard_range
    for n in (0, 1, 7, 4096, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(Exception, match="Invalid rank"):
        shard_range(10, 2, 2)


def test_reduce_is_a_noop_without_process_group():
    import torch
    from ofighters_b200.sharding import reduce_episode_stats, stats_dict
    s = torch.arange(6, dtype=torch.int64)
    assert reduce_episode_stats(s) is None and stats_dict(s)["arenas"] == 5
    with pytest.raises(Exception, match="Invalid statistics tensor"):
        reduce_episode_stats(torch.zeros(5, dtype=torch.int64))


def test_world_size_2_gloo_matches_single_process(tmp_path):
    from oracle.step_c import ArenasC
    out = str(tmp_path / "shard")
    port = str(29500 + os.getpid() % 1000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", OFB_PORT=port, OFB_OUT=out, OFB_ROOT=ROOT,
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    parts = [np.load(out + ".%d.npz" % r) for r in range(2)]
    # single-process reference over all 37 arenas
    N, S, T, seed = 37, 7, 60, 0xBEEF
    c0 = ArenasC(np.zeros((N, S, 2), np.int32))
    c = ArenasC(c0.random_spawn(seed, 0))
    for t in range(T):
        c.step(c.bot_actions("random", seed, t))
    c.reset(c.random_spawn(seed, 1))
    assert [int(p["lo"]) for p in parts] == [0, 19] and int(parts[1]["hi"]) == N
    for p in parts:                                      # weights broadcast from rank 0; [sum loss, replays] summed
        assert np.array_equal(p["flat"], np.full(10, 3.0, np.float32)) and np.array_equal(p["ls"], np.array([4.5, 3.0]))
    for k in ("x", "y", "score"):
        whole = {"x": "ship_x", "y": "ship_y", "score": "ship_reward"}[k]
        assert np.array_equal(np.concatenate([p[k] for p in parts]), c.arr[whole]), k
    # the all-reduced statistics equal the host-side sum of the local ones and the single-process totals
    assert np.array_equal(parts[0]["total"], parts[1]["total"])
    assert np.array_equal(parts[0]["total"], parts[0]["local"] + parts[1]["local"])
    assert np.array_equal(parts[0]["total"], c.arr["stats"])
