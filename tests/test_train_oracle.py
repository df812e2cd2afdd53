"""CPU checks of the Q-learning update oracle (oracle/policy_train_torch.py) and of the host-side pieces of
ofighters_b200.trainer that need no GPU (weight flattening, epsilon schedules, loud failure without CUDA)."""
import math
import os
import re

import pytest
import torch

from oracle import policy_torch as po
from oracle import policy_train_torch as pt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_batch(B=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    img = (torch.rand((B, 400, 400, 2), generator=g) < 0.02).float()
    vec = torch.rand((B, 8), generator=g) * 400
    ta = torch.randn((B, 2), generator=g)
    tp = torch.randn((B, 400, 400), generator=g) * 0.1
    return img, vec, ta, tp


def test_flat_weight_order_matches_the_header():
    from ofighters_b200.policy import WEIGHT_SPEC
    from ofighters_b200 import trainer
    n = trainer.len_flat()
    hdr = open(os.path.join(ROOT, "include", "ofb_train.h")).read()
    assert n == int(re.search(r"#define OFB_TRAIN_N_PARAMS (\d+)", hdr).group(1)) == 571730
    w = po.init_weights(3, randomize_bn=True)
    flat = trainer.flatten_weights(w)
    assert flat.numel() == n
    back = trainer.unflatten_weights(flat)
    assert all(torch.equal(back[k], w[k]) for k, _ in WEIGHT_SPEC)
    assert torch.equal(pt.flatten_weights(w, WEIGHT_SPEC), flat)
    # Appendix B parameter count: 571 730 minus the 2 x 52 moving statistics = trainable weights
    assert sum(w[k].numel() for k in pt.trainable_names(w)) == 571730 - 2 * (8 * 4 + 2 + 4 + 8)


def test_training_forward_equals_inference_forward_when_moving_stats_equal_batch_stats():
    """Pins forward_train to the already-pinned inference restatement: feed the batch statistics back as the moving ones."""
    w = po.init_weights(1, randomize_bn=True)
    img, vec, _, _ = _small_batch()
    act, ptr, stats = pt.forward_train(w, img, vec)
    w2 = dict(w)
    for bn, (mean, var, _) in stats.items():
        w2[bn + "/mean"], w2[bn + "/var"] = mean, var
    act2, ptr2 = po.forward(w2, img, vec)
    assert torch.allclose(act, act2, rtol=1e-4, atol=1e-5) and torch.allclose(ptr, ptr2, rtol=1e-4, atol=1e-5)


def test_keras_adam_first_steps_by_hand():
    opt = pt.KerasAdam(lr=1e-4)
    w = {"a": torch.tensor([1.0, -2.0, 0.5])}
    g = {"a": torch.tensor([0.3, -4.0, 0.0])}
    opt.step(w, g)
    lr_t = 1e-4 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = torch.tensor([1.0, -2.0, 0.5]) - lr_t * (0.1 * g["a"]) / (torch.sqrt(0.001 * g["a"] ** 2) + 1e-7)
    assert torch.allclose(w["a"], want, rtol=0, atol=1e-9)
    assert abs(float(w["a"][0]) - (1.0 - 1e-4)) < 1e-7 and float(w["a"][2]) == 0.5      # ~ lr * sign(g); zero grad: no move


def test_fit_descends_and_updates_moving_statistics():
    w = po.init_weights(2)
    img, vec, ta, tp = _small_batch()
    opt = pt.KerasAdam(lr=1e-3)
    losses = [pt.fit(w, opt, img, vec, ta, tp)[0] for _ in range(4)]
    assert losses[-1] < losses[0]
    assert not torch.equal(w["norm1/mean"], torch.zeros(8)) and float(w["norm1/var"].max()) < 1.0
    # gradient of one weight against a central finite difference of the loss (float64 copy of the model)
    wd = {k: v.double() for k, v in po.init_weights(2).items()}
    (_, _, _), grads, _ = pt.loss_and_grads(wd, img.double(), vec.double(), ta.double(), tp.double())
    for name, idx in (("upconv4/kernel", (1, 1, 3, 0)), ("dense2/bias", (7,)), ("norm2/gamma", (5,))):
        eps = 1e-5
        vals = []
        for s in (+1, -1):
            w2 = {k: v.clone() for k, v in wd.items()}
            w2[name][idx] += s * eps
            a, p, _ = pt.forward_train(w2, img.double(), vec.double())
            vals.append(float(((a - ta.double()) ** 2).mean() + ((p - tp.double()) ** 2).mean()))
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - float(grads[name][idx])) <= 1e-5 + 1e-3 * abs(fd), (name, fd, float(grads[name][idx]))


def test_td_targets_follow_replay_including_its_transposed_pointer_index():
    g = torch.Generator().manual_seed(5)
    B = 3
    act_o, act_n = torch.randn((B, 2), generator=g), torch.randn((B, 2), generator=g)
    ptr_o, ptr_n = torch.randn((B, 400, 400), generator=g), torch.randn((B, 400, 400), generator=g)
    ia = torch.tensor([0, 1, 1])
    pointer = torch.tensor([[10, 20], [399, 0], [7, 7]])
    reward = torch.tensor([2.0, 0.0, 3.0])
    done = torch.tensor([0, 1, 0])
    ta, tp = pt.td_targets(act_o, ptr_o, act_n, ptr_n, ia, pointer, reward, done)
    assert float(ta[0, 0]) == pytest.approx(2.0 + 0.9 * float(act_n[0].max())) and float(ta[0, 1]) == float(act_o[0, 1])
    assert float(ta[1, 1]) == 0.0                                       # done: no bootstrap
    assert float(tp[0, 10, 20]) == pytest.approx(2.0 + 0.9 * float(ptr_n[0].max()))   # [x][y], not [y][x]
    assert float(tp[0, 20, 10]) == float(ptr_o[0, 20, 10])
    assert int((tp != ptr_o).sum()) == 3


def test_epsilon_schedules_follow_lib_epsilon():
    """EpsilonSchedule's closed forms against the recurrences of lib/epsilon.py:36-86, stated here independently:
    cosine  eps_k = (cos(2 pi (k mod P) / P) + 1) / 2;  decay  eps_{k+1} = eps_k * 0.9999 while eps_k > 0.01."""
    import math
    from ofighters_b200.trainer import Epsilon_cos, Epsilon_decay
    e = Epsilon_cos(period=200)
    assert e.get() == 1.0
    vals = [e.next() for _ in range(200)]
    for k in (1, 50, 99, 100, 150, 199):
        assert vals[k - 1] == pytest.approx((math.cos(2 * math.pi * k / 200) + 1) / 2, abs=1e-12)
    assert vals[99] == pytest.approx(0.0, abs=1e-12) and vals[199] == 1.0 and e.t == 0      # t wraps at the period
    e.set(0.5)
    assert e.t == pytest.approx(50.0) and e.get() == pytest.approx(0.5)                     # reverse_cos_P_u
    d = Epsilon_decay()
    ref = 1.0
    for k in range(1, 60000):
        if ref > 0.01:
            ref *= 0.9999
        if k in (1, 3, 1000, 46049, 46052, 59999):
            assert d.value_at(k) == pytest.approx(ref, rel=1e-9), k
    for _ in range(3):
        d.next()
    assert d.get() == pytest.approx(0.9999 ** 3)
    d.set(0.005)
    assert d.next() == 0.005                                            # below the soft minimum: left alone
    with pytest.raises(Exception, match=r"\[0, 1\]"):
        d.set(1.5)


def test_epsilon_device_formula_matches_python():
    """ofb_eps_value runs the SAME eps_at() the action kernel evaluates on the device (ofb_policy.cu)."""
    import ctypes as C
    from ofighters_b200 import _lib
    from ofighters_b200.epsilon import EpsilonSchedule
    lib = _lib.load()
    for sch in (EpsilonSchedule.cosine(110 * 400), EpsilonSchedule.cosine(200, 0.8), EpsilonSchedule.decaying(),
                EpsilonSchedule.decaying(0.5, 0.999, 0.05), EpsilonSchedule.constant(0.25)):
        st = sch.as_struct()
        for t in (0, 1, 17, 199, 200, 4400, 21999.5, 46051, 46052, 10 ** 6):
            assert float(lib.ofb_eps_value(C.byref(st), float(t))) == pytest.approx(sch.value_at(t), abs=2e-7), (sch.kind, t)


def test_trainer_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ofighters_b200 import OfbError
    from ofighters_b200.trainer import TrainerB200
    with pytest.raises(OfbError, match="no CPU fallback"):
        TrainerB200()
