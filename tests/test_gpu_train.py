"""Parity of the Q-learning update kernels (include/ofb_train.h) with the torch-autograd oracle -- runs on the B200 box.

Tolerances (fp32 on both sides, different summation orders; sums over up to 1.28 M elements):
  * training-mode predictions and losses: 2e-4 relative to the tensor's max;
  * gradients, per tensor, against the oracle run in fp64: max |difference| <= 2e-3 * max |gradient|;
  * weights after Adam steps: the first steps move every weight by ~lr * sign(g), so elements whose gradient is below
    the noise of the sum may step the other way -- the update must match within 5 % of lr on at least 99 % of every
    tensor's elements and never differ by more than 2.2 * lr * steps;
  * BatchNormalization moving statistics: 1e-5 absolute.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _arena_batch(B, frames=35, seed=11):
    from ofighters_b200 import BatchedBattleground
    bg = BatchedBattleground(B, ships={"random": 7}, seed=seed)
    for _ in range(frames):
        bg.frame()
    maps = bg.raster("bits")
    vec = bg.obs_vec[:, 0, :].contiguous()
    return bg, maps, vec


def _dense_image(maps):
    b = maps.cpu().numpy().view(np.uint32)
    n = b.shape[0]
    img = np.unpackbits(b.view(np.uint8).reshape(n, 2, -1), axis=2, bitorder="little").reshape(n, 2, 400, 400)
    return torch.from_numpy(img.transpose(0, 2, 3, 1).astype(np.float32))


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("B", [8, 3])
def test_fit_matches_the_autograd_oracle(B):
    from oracle import policy_torch as po
    from oracle import policy_train_torch as pt
    from ofighters_b200.trainer import TrainerB200
    w = po.init_weights(4, randomize_bn=True)
    _, maps, vec = _arena_batch(B)
    g = torch.Generator().manual_seed(B)
    ta = torch.randn((B, 2), generator=g) * 3
    tp = torch.randn((B, 400, 400), generator=g) * 0.5
    tr = TrainerB200(weights=w, learning_rate=1e-4, batch_size=8)
    img, vec_h = _dense_image(maps), vec.cpu()

    # training-mode forward
    act, ptr = tr.forward_train(maps, vec)
    oact, optr, _ = pt.forward_train(w, img, vec_h)
    assert _rel(act.cpu(), oact) <= 2e-4 and _rel(ptr.cpu(), optr) <= 2e-4

    # one fit: loss, gradients, moving statistics, weights.  The checker runs in fp64: torch's own fp32 backward is the
    # noisier of the two (e.g. conv2/kernel: 8e-3 absolute vs 5e-5 for the kernels here, scripts/train_debug.py).
    wd = {k: v.double() for k, v in w.items()}
    imgd, vecd, tad, tpd = img.double(), vec_h.double(), ta.double(), tp.double()
    opt = pt.KerasAdam(lr=1e-4)
    (olosses, ograds, _) = pt.loss_and_grads(wd, imgd, vecd, tad, tpd)
    loss = tr.fit(maps, vec, ta.cuda(), tp.cuda()).cpu()
    pt.fit(wd, opt, imgd, vecd, tad, tpd)
    assert np.allclose(loss.numpy(), np.array(olosses), rtol=2e-4)
    grads = tr.get_grads()
    # a conv bias in front of a BatchNormalization has an analytically zero gradient (the batch mean removes it): both sides
    # hold rounding noise there, so the scale is the layer's kernel gradient
    zero_grad = [k for k in ograds if k.endswith("/bias") and "conv" in k and k != "upconv4/bias"]
    for k, og in ograds.items():
        scale = max(float(og.abs().max()), float(ograds[k.replace("/bias", "/kernel")].abs().max()) if k in zero_grad else 0.0)
        err = float((grads[k].double() - og).abs().max())
        assert err <= 2e-3 * scale + 1e-9, (k, err, scale)
        if k in zero_grad:
            assert float(og.abs().max()) <= 1e-9 and float(grads[k].abs().max()) <= 1e-3 * scale
    for steps in (1, 3):
        if steps == 3:
            for _ in range(2):
                tr.fit(maps, vec, ta.cuda(), tp.cuda())
                pt.fit(wd, opt, imgd, vecd, tad, tpd)
        got = tr.get_weights()
        assert tr.steps == steps
        for k in wd:
            if k.endswith("/mean") or k.endswith("/var"):
                assert float((got[k].double() - wd[k]).abs().max()) <= 1e-5, k
                continue
            diff = ((got[k].double() - w[k].double()) - (wd[k] - w[k].double())).abs()
            assert float(diff.max()) <= 2.2e-4 * steps, (k, float(diff.max()))
            if k not in zero_grad:                       # (there Adam turns the rounding noise into +-lr steps on either side)
                assert float((diff <= 5e-6 * steps).float().mean()) >= 0.99, (k, float((diff <= 5e-6 * steps).float().mean()))
    got = {k: v.float() for k, v in got.items()}
    # the inference engine was refreshed with the trained weights (BN folded from the moving statistics)
    r = tr.model.forward(maps, vec, 1, want_ptr=True)
    iact, iptr = po.forward(got, img, vec_h)
    assert _rel(r["act"].cpu(), iact) <= 2.5e-2 and _rel(r["ptr"].cpu(), iptr) <= 2.5e-2


def test_td_targets_kernel_is_exact_and_replay_runs():
    import random
    from oracle import policy_train_torch as pt
    from ofighters_b200 import _lib
    from ofighters_b200.trainer import Epsilon_cos, TrainerB200
    import ctypes as C
    B = 5
    g = torch.Generator().manual_seed(9)
    act_o, act_n = torch.randn((B, 2), generator=g), torch.randn((B, 2), generator=g)
    ptr_o, ptr_n = torch.randn((B, 400, 400), generator=g), torch.randn((B, 400, 400), generator=g)
    ia = torch.tensor([0, 1, 1, 0, 1], dtype=torch.int32)
    pointer = torch.tensor([[10, 20], [399, 0], [7, 7], [0, 399], [200, 123]], dtype=torch.int32)
    reward = torch.tensor([2.0, 0.0, 3.0, 1.0, 0.0])
    done = torch.tensor([0, 1, 0, 0, 1], dtype=torch.uint8)
    want_a, want_p = pt.td_targets(act_o, ptr_o, act_n, ptr_n, ia, pointer, reward, done)
    d = [t.cuda() for t in (act_o, ptr_o, act_n, ptr_n, ia, pointer, reward, done)]
    ta, tp = torch.empty_like(d[0]), torch.empty_like(d[1])
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.load().ofb_trainer_td_targets(*[p(t) for t in d], 0.9, B, p(ta), p(tp), None))
    torch.cuda.synchronize()
    assert torch.equal(ta.cpu(), want_a) and torch.equal(tp.cpu(), want_p)

    # Trainer surface: remember / replay / get_best_action on observations of a running arena batch
    random.seed(3)
    bg, maps, vec = _arena_batch(16, frames=10, seed=5)
    tr = TrainerB200(learning_rate=1e-4, epsilon=Epsilon_cos(period=110 * 400), batch_size=8, memory_size=400)
    prev = None
    for t in range(12):
        maps = bg.raster("bits")
        obs = [(maps[k].clone(), bg.obs_vec[k, 0].clone()) for k in range(16)]
        if prev is not None:
            for k in range(16):
                tr.remember(prev[k][0], prev[k][1], prev[k][2], float(bg.obs_vec[k, 0, 0]), obs[k], False)
        acts = [tr.random_play() for _ in range(16)]
        prev = [(obs[k], acts[k][0], acts[k][1]) for k in range(16)]
        bg.frame()
    assert len(tr.memory) == 11 * 16
    w0 = tr.get_weights()
    hist = tr.replay(tr.batch_size)
    assert np.isfinite(hist.history["loss"][0]) and tr.steps == 1
    w1 = tr.get_weights()
    assert any(not torch.equal(w0[k], w1[k]) for k in w0)
    tr.epsilon.set(0.0)
    ia_, (x, y) = tr.get_best_action(maps[0], bg.obs_vec[0, 0], rand=True)
    assert ia_ in (0, 1) and 0 <= x < 400 and 0 <= y < 400
    k = int(torch.argmax(tr.ptr_values.reshape(-1)))
    assert (x, y) == (k % 400, k // 400)
    with pytest.raises(Exception, match="Invalid batch"):
        tr.fit(torch.zeros((9, 2, 5000), dtype=torch.int32, device="cuda"), torch.zeros((9, 8)), torch.zeros((9, 2)),
               torch.zeros((9, 400, 400)))


def test_qlearner_follows_qlearnia_bookkeeping():
    """QlearnIA.play's bookkeeping on a batch (agents/qlearnIA_V2.py:372-417): one transition per live tracked policy ship
    and frame, none after the ship has seen its own death; a replay at the death of ANY tracked bot and, while bot 1 is
    still playing, every 50 total steps; epsilon advanced once per decision of bot 1 outside the 20-step collecting phase;
    previous_* forgotten at a restart; a snapshot at the first frame of every `snapshot`-th episode; and the inference
    engine tracks the trained weights."""
    import random
    import tempfile
    from ofighters_b200 import BatchedBattleground
    from ofighters_b200.trainer import Epsilon_cos, QLearner, TrainerB200
    random.seed(1)
    bg = BatchedBattleground(32, ships={"QlearnIA": 1, "random": 6}, seed=21)
    maps = bg.raster("bits")
    tr = TrainerB200(learning_rate=1e-4, epsilon=Epsilon_cos(period=110 * 400), batch_size=8, max_ships=32, name="t")
    folder = tempfile.mkdtemp()
    ql = QLearner(tr, track=4, replay_every=50, collecting_steps=20, snapshot=1, snapshot_folder=folder)
    w0 = tr.get_weights()
    expect, alive_prev, had_prev = 0, [True] * 4, [False] * 4
    exp_replays, exp_eps_t = 0, 0
    for t in range(120):
        if t == 100:
            bg.restart()
            bg.raster("bits", out=maps)
            ql.reset()
            alive_prev, had_prev = [True] * 4, [False] * 4
        assert ql.collecting == (t + 1 < 20)
        alive = bg.state(("ship_alive",))["ship_alive"][:4, 0].tolist()
        mem_nonempty = len(tr.memory) > 0
        for k in range(4):
            if alive_prev[k]:                            # the bot had not yet seen its death before this frame
                if not alive[k] and (mem_nonempty or expect > 0):
                    exp_replays += 1                     # replay on its own death (before this frame's remember)
                expect += 1 if had_prev[k] else 0
                had_prev[k] = True
                if k == 0:
                    exp_eps_t += 0 if t + 1 < 20 else 1
                    if (t + 1) % 50 == 0:
                        exp_replays += 1
            alive_prev[k] = alive_prev[k] and bool(alive[k])
        iact, xy, _ = ql.act(bg, maps)
        assert min(len(tr.memory), 400) == min(expect, 400), (t, len(tr.memory), expect)
        # the remembered previous action is the row that was written (eps-greedy / collecting included)
        row = bg.actions[0, 0].tolist()
        if ql.previous[0] is not None and alive[0]:
            _, pa, pp = ql.previous[0]
            assert row == [1 if pa == 0 else 0, 1 if pa == 1 else 0, pp[0], pp[1]], (t, row, pa, pp)
        bg.frame()
        bg.raster("bits", out=maps)
    assert ql.total_steps == 120 and tr.steps == exp_replays and len(ql.losses) == tr.steps and all(np.isfinite(ql.losses))
    assert tr.steps >= 1
    w1 = tr.get_weights()
    assert not torch.equal(w0["upconv4/kernel"], w1["upconv4/kernel"])
    assert torch.equal(tr.model.weights["dense2/kernel"], w1["dense2/kernel"])      # inference engine refreshed after each fit
    assert tr.epsilon.t == exp_eps_t                     # decay_epsilon once per decision of bot 1 (:398-401)
    assert len(ql.saved) == 1 and ql.epsilons and ql.episode == 1
    back = TrainerB200.load_weights_file(ql.saved[0])
    assert set(back) == set(w1) and all(back[k].shape == w1[k].shape for k in w1)


def test_remembered_action_is_the_row_played():
    """with epsilon = 1 every policy row is random_play(); the (iaction, pointer) handed back by act() -- what
    QLearner remembers -- must be exactly what landed in bg.actions, and differ from the greedy prediction."""
    from ofighters_b200 import BatchedBattleground
    from ofighters_b200.policy import PolicyB200
    from ofighters_b200.epsilon import EpsilonSchedule
    bg = BatchedBattleground(256, ships={"QlearnIA": 1, "random": 6}, seed=5)
    pol = PolicyB200.random_init(max_ships=256)
    maps = bg.raster("bits")
    gi, gxy = pol.decide(bg, maps)
    gi, gxy = gi.clone(), gxy.clone()
    for eps, collecting in ((EpsilonSchedule.constant(1.0), False), (1.0, False), (0.0, True)):
        iact, xy = pol.act(bg, maps, epsilon=eps, collecting=collecting)
        rows = bg.actions[:, 0, :].cpu().long()
        ia, p = iact.cpu().long(), xy.cpu().long()
        assert torch.equal(rows[:, 0], (ia == 0).long()) and torch.equal(rows[:, 1], (ia == 1).long())
        assert torch.equal(rows[:, 2:], p)
        assert int(p.min()) >= 0 and int(p.max()) <= 399 and set(ia.tolist()) == {0, 1}
        assert not torch.equal(p, gxy.cpu().long())      # 256 uniform draws do not reproduce the prediction
    iact, xy = pol.act(bg, maps, epsilon=0.0)            # greedy: untouched
    assert torch.equal(iact, gi) and torch.equal(xy, gxy)
    half = EpsilonSchedule.cosine(400)
    half.set(0.5)
    iact, xy = pol.act(bg, maps, epsilon=half)
    changed = float(((xy != gxy).any(dim=1)).float().mean())
    assert 0.3 < changed < 0.7, changed                  # about half the rows explore at eps(t) = 0.5


def test_viewer_bridge_streams_one_arena():
    """what the Tk controller and ActionMapGraph read, pulled from one arena of a GPU-stepped batch."""
    from ofighters_b200 import BatchedBattleground
    from ofighters_b200.trainer import TrainerB200
    from ofighters_b200.viewer import ViewerBridge
    bg = BatchedBattleground(24, ships={"QlearnIA": 1, "random": 6}, seed=33)
    tr = TrainerB200(learning_rate=1e-4, batch_size=8, max_ships=32)
    maps = bg.raster("bits")
    for _ in range(25):
        tr.model.act(bg, maps)
        bg.frame(maps=maps)
    v = ViewerBridge(bg, arena=5, trainer=tr).refresh(maps)
    st = {k: t[5].cpu().numpy() for k, t in bg.state().items()}
    assert [s.body.x for s in v.battleground.ships] == st["ship_x"].tolist()
    assert len(v.battleground.lasers) == int(st["n_lasers"])
    # maps: [row = y, col = x] like Observation.ship_map; every live ship's centre pixel is set
    assert v.ship_map.shape == (400, 400) and v.laser_map.shape == (400, 400)
    for s in v.battleground.ships:
        if s.state == "flying" and 0 <= s.body.x < 400 and 0 <= s.body.y < 400:
            assert v.ship_map[s.body.y, s.body.x] == 1.0
    assert int(v.ship_map.sum()) == int(np.unpackbits(maps[5, 0].cpu().numpy().view(np.uint8)).sum())
    # ActionMapGraph's inputs: the attached trainer's act_values (2,) / ptr_values (400, 400) of this arena's policy ship
    assert tr.act_values.shape == (2,) and tr.ptr_values.shape == (400, 400)
    r = tr.model.forward(maps[5:6].contiguous(), bg.obs_vec[5, 0].reshape(1, 8), 1, want_ptr=True)
    assert np.array_equal(v.ptr_values, r["ptr"][0].cpu().numpy())
    o = v.observation()
    assert o["pos"] == (v.battleground.ships[0].body.x, v.battleground.ships[0].body.y) and o["dim"] == (400, 400)
    with pytest.raises(Exception, match="Invalid arena"):
        ViewerBridge(bg, arena=24)
    # ScoreGraph / EpsilonGraph inputs: agent.scores gets the finished episode's score at every restart (agents/agent.py:59-64),
    # QlearnIA.epsilons the trainer's epsilon at every reset (agents/qlearnIA_V2.py:364)
    from ofighters_b200.trainer import QLearner
    ql = QLearner(tr, track=2, snapshot=0)
    v2 = ViewerBridge(bg, arena=5, trainer=tr, learner=ql)
    assert v2.scores == [] and v2.epsilons == []
    want = int(bg.state(("ship_score",))["ship_score"][5, 0])
    bg.restart()
    ql.reset()
    assert v2.scores == [want] and v.scores == [want] and v2.epsilons == [tr.epsilon.get()] and v2.losses == ql.losses   # one list per watched arena


def test_replay_is_the_references_sequence_of_calls():
    """Trainer.replay (:240-287) spelled out call by call on a twin trainer: sample the minibatch with the same ``random``
    state, predict on obs and next_obs, build the targets with the oracle's td_targets, then fit on the NEXT observations
    (``inputs1[i] = img_input`` is evaluated after img_input was rebuilt from next_obs) -- the two trainers must end up with the
    same weights (fp32 atomics aside) and the same reported loss."""
    import random
    from oracle import policy_torch as po
    from oracle import policy_train_torch as pt
    from ofighters_b200.trainer import TrainerB200
    w = po.init_weights(6, randomize_bn=True)
    bg, _, _ = _arena_batch(12, frames=8, seed=2)
    a, b = TrainerB200(weights=w, learning_rate=1e-4, batch_size=8), TrainerB200(weights=w, learning_rate=1e-4, batch_size=8)
    g = torch.Generator().manual_seed(4)
    prev = None
    for t in range(6):
        maps = bg.raster("bits")
        obs = [(maps[k].clone(), bg.obs_vec[k, 0].clone()) for k in range(12)]
        if prev is not None:
            for k in range(12):
                tr_args = (prev[k][0], prev[k][1], prev[k][2], float(bg.obs_vec[k, 0, 0]), obs[k], bool(t == 5 and k % 3 == 0))
                a.remember(*tr_args)
                b.remember(*tr_args)
        prev = [(obs[k], int(torch.randint(0, 2, (1,), generator=g)), (int(torch.randint(0, 400, (1,), generator=g)),
                                                                         int(torch.randint(0, 400, (1,), generator=g)))) for k in range(12)]
        bg.frame()
    random.seed(7)
    hist = a.replay(8)
    # the same thing by hand on the twin
    random.seed(7)
    mb = random.sample(b.memory, 8)
    maps_o, vec_o = torch.stack([m[0][0] for m in mb]), torch.stack([m[0][1] for m in mb])
    maps_n, vec_n = torch.stack([m[4][0] for m in mb]), torch.stack([m[4][1] for m in mb])
    act_o, ptr_o = b.predict(maps_o.contiguous(), vec_o)
    act_n, ptr_n = b.predict(maps_n.contiguous(), vec_n)
    ta, tp = pt.td_targets(act_o.cpu(), ptr_o.cpu(), act_n.cpu(), ptr_n.cpu(), torch.tensor([m[1] for m in mb]),
                           torch.tensor([list(m[2]) for m in mb]), torch.tensor([m[3] for m in mb], dtype=torch.float32),
                           torch.tensor([1 if m[5] else 0 for m in mb]))
    loss = b.fit(maps_n.contiguous(), vec_n, ta.cuda(), tp.cuda()).cpu()
    assert hist.history["loss"][0] == pytest.approx(float(loss[0]), rel=1e-5)
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        # The weight gradients add with fp32 atomics, and Adam's first step is lr * g / (|g| + 1e-7): an element whose gradient
        # is at the level of that rounding noise (and every conv bias in front of a BatchNormalization, whose gradient is
        # analytically zero) may step differently on the two trainers -- by at most 2 * lr; everything else must agree.
        noise_driven = k.endswith("/bias") and "conv" in k and k != "upconv4/bias"
        diff = (wa[k] - wb[k]).abs()
        assert float(diff.max()) <= 2.2e-4, k
        if not noise_driven:
            assert float((diff <= 2e-6 + 1e-5 * float(wb[k].abs().max())).float().mean()) >= 0.999, k


def test_player_bridge_is_read_keys():
    """Human Player passthrough (lib/player.py, Ship.read_keys): pending keys -> the external ship's action row."""
    from ofighters_b200 import BatchedBattleground
    from ofighters_b200.viewer import PlayerBridge
    from types import SimpleNamespace
    bg = BatchedBattleground(6, ships={"external": 1, "random": 6}, seed=8)
    pl = PlayerBridge(bg, arena=2, ship=0)
    x0, y0 = [int(v) for v in bg.obs_vec[2, 0, 6:8].tolist()]
    row = pl.write_action().tolist()                     # nothing pressed: no shot, no thrust, pointing unchanged (= own position)
    assert row == [0, 0, x0, y0]
    pl.press_shoot()
    pl.request_thrust()
    pl.request_turn(SimpleNamespace(x=123, y=45))
    assert pl.write_action().tolist() == [1, 1, 123, 45]
    bg.frame()                                           # the six device bots draw their own rows; row (2, 0) is the player's
    assert [int(v) for v in bg.obs_vec[2, 0, 2:4].tolist()] == [123, 45]
    assert int(bg.state(("shots",))["shots"][2]) >= 1
    assert pl.write_action().tolist() == [1, 0, 123, 45]     # the button is still held, thrust was a key press, cursor did not move
    pl.unpress_shoot()
    assert pl.write_action().tolist() == [0, 0, 123, 45]
    with pytest.raises(Exception, match="device bot"):
        PlayerBridge(bg, arena=0, ship=3)


def test_repeated_fit_descends_on_the_device():
    """Ten Adam steps on one fixed batch: the loss the kernels report goes down (and keeps tracking the oracle's trajectory)."""
    from oracle import policy_torch as po
    from oracle import policy_train_torch as pt
    from ofighters_b200.trainer import TrainerB200
    w = po.init_weights(9)
    _, maps, vec = _arena_batch(4, frames=20, seed=31)
    g = torch.Generator().manual_seed(2)
    ta, tp = torch.randn((4, 2), generator=g), torch.randn((4, 400, 400), generator=g) * 0.3
    tr = TrainerB200(weights=w, learning_rate=1e-3, batch_size=4, max_ships=8)
    img, vec_h = _dense_image(maps), vec.cpu()
    wo, opt = {k: v.clone() for k, v in w.items()}, pt.KerasAdam(lr=1e-3)
    losses, olosses = [], []
    for _ in range(10):
        losses.append(float(tr.fit(maps, vec, ta.cuda(), tp.cuda(), sync_model=False)[0]))
        olosses.append(pt.fit(wo, opt, img, vec_h, ta, tp)[0])
    assert losses[-1] < 0.8 * losses[0] and all(np.isfinite(losses))
    assert np.allclose(losses, olosses, rtol=2e-2), (losses, olosses)      # two fp32 trajectories of a 10-step descent
