"""BASELINE.json's full sizes on the B200 box, checked through size-independent properties (the oracles finish such sizes
in minutes, not seconds): arenas are independent and keyed by global id, so any window of a full-size batch must equal a
small batch of the same global ids (which the other GPU tests pin to the oracle); discrete counters obey the reference's
bookkeeping identities; the fused frame kernel equals the two-launch form; the chunked policy forward equals the forward
of a window."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("time", "n_lasers", "kills", "deaths", "shots", "overflow", "ship_x", "ship_y", "ship_px", "ship_py", "ship_alive",
        "ship_hull", "ship_reward", "ship_score", "ship_steps")


def _window_equal(big, small, lo, hi, maps_big, maps_small, tag):
    sb, ss = big.state(), small.state()
    for k in KEYS:
        assert torch.equal(sb[k][lo:hi], ss[k]), (tag, k)
    live = torch.arange(big.laser_cap, device=big.device)[None, :] < ss["n_lasers"].long()[:, None]
    for k in ("laser_x", "laser_y", "laser_owner", "laser_destroyed"):
        assert torch.equal(torch.where(live, sb[k][lo:hi], torch.zeros_like(ss[k])),
                           torch.where(live, ss[k], torch.zeros_like(ss[k]))), (tag, k)
    assert torch.equal(big.obs_vec[lo:hi], small.obs_vec), tag
    assert torch.equal(maps_big[lo:hi], maps_small), tag


@pytest.mark.parametrize("N,S,kind,lcap,T", [(65536, 7, "random", 0, 230), (131072, 7, "random", 0, 40),
                                             (16384, 32, "stress", 2048, 30)],
                         ids=["configs2-65536", "configs4-131072-per-gpu", "configs3-stress-16384x32"])
def test_full_size_batch_equals_small_batches_and_obeys_the_bookkeeping(N, S, kind, lcap, T):
    from ofighters_b200 import ArenaConfig, BatchedBattleground
    cfg = ArenaConfig(laser_cap=lcap)
    big = BatchedBattleground(N, ships={kind: S}, config=cfg, seed=0x0F16, arena0=5)
    windows = [(0, 96), (N // 2 - 13, N // 2 + 51), (N - 80, N)]
    smalls = [BatchedBattleground(hi - lo, ships={kind: S}, config=cfg, seed=0x0F16, arena0=5 + lo) for lo, hi in windows]
    maps = big.raster("bits")
    smaps = [s.raster("bits") for s in smalls]
    for t in range(T):
        if t == 200:                                     # MAX_TIME: the episode restarts (lib/ofighters.py:684-688)
            for b in [big] + smalls:
                b.restart()
        big.frame(maps=maps)                             # fused frame kernel at full size
        for s, m in zip(smalls, smaps):
            s.frame()                                    # two-launch form on the windows
            s.raster("bits", out=m)
    for (lo, hi), s, m in zip(windows, smalls, smaps):
        _window_equal(big, s, lo, hi, maps, m, "window %d:%d" % (lo, hi))
    st = big.state()
    # hull is 1 and never restored: a ship dies at most once per episode, every kill is a death (lib/ship.py:127-131,225-230)
    dead = (S - st["ship_alive"].long().sum(dim=1))
    assert torch.equal(st["kills"], st["deaths"]) and torch.equal(st["deaths"].long(), dead)
    assert int(st["overflow"].sum()) == 0 and int(st["n_lasers"].max()) <= big.laser_cap
    assert int((st["time"] != (T - 200 if T > 200 else T)).sum()) == 0
    # the laser list only holds lasers that were shot this episode
    assert bool((st["n_lasers"] <= st["shots"]).all())
    # checksum of checksums: the ship map holds at most one 193-pixel disk per live ship, the laser map at most 13 pixels per laser
    pop = lambda x: int(np.unpackbits(x.cpu().numpy().view(np.uint8)).sum())
    sample = slice(1000, 1256)
    assert pop(maps[sample, 0]) <= 193 * int(st["ship_alive"][sample].sum())
    assert pop(maps[sample, 1]) <= 13 * int(st["n_lasers"][sample].sum())
    if T > 200:                                          # the per-episode statistics: [sum score, kills, deaths, shots, ships, arenas]
        stats = big.stats.cpu().tolist()
        assert stats[4] == N * S and stats[5] == N and stats[1] == stats[2] and stats[3] > 0


def test_chunked_policy_forward_at_configs2_size_equals_window_forwards():
    """65 536 arenas through PolicyB200 in chunks of 8 192: the decoded actions of any window equal a forward of that window."""
    from ofighters_b200 import BatchedBattleground
    from ofighters_b200.policy import PolicyB200
    N = 65536
    bg = BatchedBattleground(N, ships={"QlearnIA": 1, "random": 6}, seed=0x0F16)
    maps = bg.raster("bits")
    pol = PolicyB200.random_init(seed=0, max_ships=8192)
    for t in range(12):
        pol.act(bg, maps)
        bg.frame(maps=maps)
    vec = bg.obs_vec[:, 0, :].contiguous()
    ia, xy = pol.forward_argmax(maps, vec)
    assert int(ia.min()) >= 0 and int(ia.max()) <= 1 and int(xy.min()) >= 0 and int(xy.max()) <= 399
    for lo, hi in ((0, 300), (8192 - 100, 8192 + 100), (N - 257, N)):
        ia_w, xy_w = pol.forward_argmax(maps[lo:hi].contiguous(), vec[lo:hi].contiguous())
        assert torch.equal(ia[lo:hi], ia_w) and torch.equal(xy[lo:hi], xy_w), (lo, hi)
    # the action rows written for the policy ships are QlearnIA.play's vector (qlearnIA_V2.py:447-456)
    pol.act(bg, maps)
    rows = bg.actions[:, 0, :].long()
    assert torch.equal(rows[:, 0], (ia == 0).long()) and torch.equal(rows[:, 1], (ia == 1).long())
    assert torch.equal(rows[:, 2:], xy.long())
