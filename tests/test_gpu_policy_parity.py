"""Hardened parity of the policy forward (product path: tcgen05 engine through the C ABI) with the fp32 oracle -- B200 box.

VERDICT r1 item 4: 512 arenas sampled at frames {0, 20, 40, 100, 180} of an episode plus the 32-ship stress scene, for the
Keras-default weights the bench uses AND randomised BatchNormalization statistics, one and seven policy ships per arena,
sparse and dense trunk each against the fp32 oracle DIRECTLY, both bilinear kernels.  Tolerances are the measured level
(bf16 activations between layers, bf16 tensor-core operands, fp32 accumulation) with a small margin:
    act   max|x - ref| <= ACT_TOL * max(1, max|ref|)    per batch
    ptr   max|x - ref| <= PTR_TOL * max|ref|            per ship
Discrete outputs: the decoded (iaction, x, y) must be a near-argmax of the oracle's outputs for EVERY ship (value at our
choice >= the oracle's maximum - tolerance) and the exact-agreement rates are reported and floored.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# Measured on B200 over 6 scenes x 512 ships x 2 weight sets x 2 trunks (r02): act <= 4.7e-4; ptr <= 1.3e-2 with the
# Keras-default weights and <= 2.03e-2 with randomised BN statistics (medians 7e-3 / 1.2e-2); exact agreement with the
# oracle's decode: iaction 100 %, (x, y) 66 % / 91 % (the pointer map has plateaus within bf16 noise of its maximum).
ACT_TOL = float(os.environ.get("OFB_ACT_TOL", "1e-3"))       # (the environment overrides exist to survey the error distribution)
PTR_TOL = float(os.environ.get("OFB_PTR_TOL", "2.5e-2"))
FRAMES = (0, 20, 40, 100, 180)
N_ARENAS = 512


def _dense_image(maps):
    b = maps.cpu().numpy().view(np.uint32)
    d = np.unpackbits(b.view(np.uint8).reshape(b.shape[0], 2, -1), axis=2, bitorder="little")
    return torch.from_numpy(d.reshape(b.shape[0], 2, 400, 400).transpose(0, 2, 3, 1).astype(np.float32))


def _oracle(w, maps, vec, P, bilinear="tf2"):
    """fp32 oracle on every (arena, policy ship): the trunk is recomputed per ship exactly like the reference's batch-1 predict."""
    from oracle import policy_torch as po
    img = _dense_image(maps)
    if P > 1:
        img = img.repeat_interleave(P, dim=0)
    acts, ptrs = [], []
    with torch.no_grad():
        for i in range(0, img.shape[0], 64):
            a, p = po.forward(w, img[i:i + 64], vec[i:i + 64].cpu(), bilinear=bilinear)
            acts.append(a)
            ptrs.append(p)
    return torch.cat(acts), torch.cat(ptrs)


@pytest.fixture(scope="module")
def scenes():
    """Bit maps + observation heads of 512 default arenas at five points of an episode, and of 64 stress arenas (32 ships,
    every ship shooting) at frame 12."""
    from ofighters_b200 import ArenaConfig, BatchedBattleground
    out = {}
    bg = BatchedBattleground(N_ARENAS, ships={"random": 7}, seed=77)
    maps = bg.raster("bits")
    t = 0
    for f in FRAMES:
        while t < f:
            bg.frame(maps=maps)
            t += 1
        out["frame%d" % f] = (maps.clone(), bg.obs_vec.clone())
    st = BatchedBattleground(64, ships={"stress": 32}, config=ArenaConfig(laser_cap=32 * 64), seed=78)
    smaps = st.raster("bits")
    for _ in range(12):
        st.frame(maps=smaps)
    out["stress"] = (smaps.clone(), st.obs_vec.clone())
    return out


@pytest.fixture(scope="module")
def weight_sets():
    from oracle import policy_torch as po
    return {"keras_default": po.init_weights(0), "random_bn": po.init_weights(5, randomize_bn=True)}


_ORACLE_CACHE = {}


def _cached_oracle(key, fn):
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = fn()
    return _ORACLE_CACHE[key]


def _compare(r, act_ref, ptr_ref, tag):
    act, ptr = r["act"].cpu(), r["ptr"].cpu()
    ea = float((act - act_ref).abs().max() / max(1.0, float(act_ref.abs().max())))
    scale = ptr_ref.abs().amax(dim=(1, 2)).clamp_min(1e-6)
    ep_ship = (ptr - ptr_ref).abs().amax(dim=(1, 2)) / scale
    ep = float(ep_ship.max())
    # discrete outputs
    n = act.shape[0]
    ia, xy = r["iaction"].cpu().long(), r["xy"].cpu().long()
    ia_ref = torch.argmax(act_ref, dim=1)
    k_ref = torch.argmax(ptr_ref.reshape(n, -1), dim=1)
    xy_ref = torch.stack([k_ref % 400, k_ref // 400], dim=1)
    ia_rate = float((ia == ia_ref).float().mean())
    xy_rate = float((xy == xy_ref).all(dim=1).float().mean())
    # our choice is a near-argmax of the oracle's outputs
    rows = torch.arange(n)
    near_a = act_ref[rows, ia] >= act_ref.max(dim=1).values - 2 * ACT_TOL * max(1.0, float(act_ref.abs().max()))
    near_p = ptr_ref[rows, xy[:, 1], xy[:, 0]] >= ptr_ref.reshape(n, -1).max(dim=1).values - 2 * PTR_TOL * scale
    # and exactly the argmax of the library's own outputs
    own = torch.argmax(ptr.reshape(n, -1), dim=1)
    assert torch.equal(xy, torch.stack([own % 400, own // 400], dim=1)), tag
    assert torch.equal(ia, (act[:, 1] > act[:, 0]).long()), tag
    print("%-44s act %.2e ptr %.2e (median %.2e) | iaction agree %.4f xy agree %.4f" % (tag, ea, ep, float(ep_ship.median()), ia_rate, xy_rate))
    assert ea <= ACT_TOL, (tag, ea)
    assert ep <= PTR_TOL, (tag, ep)
    assert bool(near_a.all()) and bool(near_p.all()), (tag, int((~near_a).sum()), int((~near_p).sum()))
    return ia_rate, xy_rate


@pytest.mark.parametrize("trunk", ["sparse", "dense"])
@pytest.mark.parametrize("wname", ["keras_default", "random_bn"])
def test_forward_parity_over_an_episode(scenes, weight_sets, wname, trunk):
    from ofighters_b200.policy import PolicyB200
    w = weight_sets[wname]
    pol = PolicyB200(w, max_ships=N_ARENAS, dense_trunk=(trunk == "dense"))
    rates = []
    for name, (maps, obs) in scenes.items():
        vec = obs[:, 0, :].contiguous()
        act_ref, ptr_ref = _cached_oracle((wname, name, 1, "tf2"), lambda: _oracle(w, maps, vec, 1))
        r = pol.forward(maps, vec, 1, want_ptr=True)
        rates.append(_compare(r, act_ref, ptr_ref, "%s/%s/%s" % (wname, trunk, name)))
    ia = float(np.mean([x[0] for x in rates])); xy = float(np.mean([x[1] for x in rates]))
    print("%s/%s: mean exact agreement iaction %.4f, (x, y) %.4f over %d scenes" % (wname, trunk, ia, xy, len(rates)))
    assert ia >= 0.99, ia
    assert xy >= 0.5, xy                                     # bf16 noise vs plateaus of the pointer map: near-argmax holds for all


@pytest.mark.parametrize("wname", ["keras_default", "random_bn"])
def test_forward_parity_seven_policy_ships(scenes, weight_sets, wname):
    """P = 7: trunk once per arena, heads per ship (observation heads differ per ship)."""
    from ofighters_b200.policy import PolicyB200
    w = weight_sets[wname]
    maps, obs = scenes["frame40"]
    maps, vec = maps[:64].contiguous(), obs[:64].reshape(-1, 8).contiguous()
    act_ref, ptr_ref = _cached_oracle((wname, "frame40", 7, "tf2"), lambda: _oracle(w, maps, vec, 7))
    pol = PolicyB200(w, max_ships=448)
    r = pol.forward(maps, vec, 7, want_ptr=True)
    _compare(r, act_ref, ptr_ref, "%s/P7/frame40" % wname)


def test_forward_parity_tf1_legacy_bilinear(scenes, weight_sets):
    from ofighters_b200.policy import PolicyB200
    w = weight_sets["random_bn"]
    maps, obs = scenes["frame40"]
    maps, vec = maps[:128].contiguous(), obs[:128, 0, :].contiguous()
    act_ref, ptr_ref = _oracle(w, maps, vec, 1, bilinear="tf1")
    act_tf2, ptr_tf2 = _cached_oracle(("random_bn", "frame40", 1, "tf2"), lambda: _oracle(w, scenes["frame40"][0], scenes["frame40"][1][:, 0, :].contiguous(), 1))
    assert float((ptr_ref - ptr_tf2[:128]).abs().max()) > 0.05 * float(ptr_ref.abs().max())            # the two kernels really differ
    for fused in (True, False):
        pol = PolicyB200(w, max_ships=128, bilinear="tf1", fused_tail=fused)
        r = pol.forward(maps, vec, 1, want_ptr=True)
        _compare(r, act_ref, ptr_ref, "random_bn/tf1/%s" % ("fused" if fused else "two-kernel tail"))
    pol.set_engine("cuda_core")
    _compare(pol.forward(maps, vec, 1, want_ptr=True), act_ref, ptr_ref, "random_bn/tf1/cuda_core")


def test_fused_tail_equals_two_kernel_tail_decisions(scenes, weight_sets):
    """The fused tail and the two-kernel tail decode the same pointer on every ship (their maps agree to bf16 rounding of
    upconv3's border pixels); the argmax-only product path equals the path that also writes the dense map."""
    from ofighters_b200.policy import PolicyB200
    w = weight_sets["keras_default"]
    maps, obs = scenes["frame20"]
    vec = obs[:, 0, :].contiguous()
    a = PolicyB200(w, max_ships=N_ARENAS)
    b = PolicyB200(w, max_ships=N_ARENAS, fused_tail=False)
    ra, rb = a.forward(maps, vec, 1, want_ptr=True), b.forward(maps, vec, 1, want_ptr=True)
    d = float((ra["ptr"] - rb["ptr"]).abs().max() / rb["ptr"].abs().max())
    agree = float((ra["xy"] == rb["xy"]).all(dim=1).float().mean())
    print("fused vs two-kernel tail: max rel diff %.2e, same (x, y) on %.4f of the ships" % (d, agree))
    assert d <= 1.5e-2 and agree >= 0.9                     # measured 1.0e-2 / 0.92 (tf32 upconv2 vs fp32 in the two-kernel path): plateau ties
    i2, xy2 = a.forward_argmax(maps, vec, 1)
    assert torch.equal(xy2, ra["xy"]) and torch.equal(i2, ra["iaction"])


@pytest.mark.parametrize("n", [1, 5, 149])
def test_cta_pair_tail_equals_single_cta_tail(scenes, weight_sets, n):
    """The fused tail as CTA pairs (tcgen05 cta_group::2: M = 256 MMAs over two SMs, each holding half of every weight operand's
    columns) gives bit-identical pointer maps and decisions, for ship counts that leave a pair's second CTA without a ship of its
    own (odd counts, fewer ships than CTAs)."""
    from ofighters_b200.policy import PolicyB200
    w = weight_sets["random_bn"]
    maps, obs = scenes["frame40"]
    maps, vec = maps[:n].contiguous(), obs[:n, 0, :].contiguous()
    a = PolicyB200(w, max_ships=max(16, n))
    b = PolicyB200(w, max_ships=max(16, n), tail_pair=True)
    ra, rb = a.forward(maps, vec, 1, want_ptr=True), b.forward(maps, vec, 1, want_ptr=True)
    assert torch.equal(ra["ptr"], rb["ptr"]) and torch.equal(ra["xy"], rb["xy"]) and torch.equal(ra["iaction"], rb["iaction"])
    ia, xya = a.forward_argmax(maps, vec, 1)
    ib, xyb = b.forward_argmax(maps, vec, 1)
    assert torch.equal(xya, xyb) and torch.equal(ia, ib)
