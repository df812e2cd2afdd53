"""Pins for the policy oracle (oracle/policy_torch.py): the conventions SURVEY Appendix B states,
checked by hand-computable cases.  The oracle is 'parity unpinned' against Keras itself (absent)."""
import numpy as np
import torch

from oracle import policy_torch as po


def test_parameter_and_mac_counts_match_the_layer_table():
    w = po.init_weights(0)
    n = sum(v.numel() for k, v in w.items() if k.endswith("kernel") or k.endswith("bias"))
    bn = sum(v.numel() for k, v in w.items() if k.split("/")[1] in ("gamma", "beta", "mean", "var"))
    # Appendix B: 152+584*3 convs, 500900+5050+102+63125 dense, 20+76+296+73 up-convs; BN 32*4+8+16+32
    assert n == 152 + 3 * 584 + 500900 + 5050 + 102 + 63125 + 20 + 76 + 296 + 73
    assert bn == 4 * 32 + 8 + 16 + 32
    macs = (400 * 400 * 18 * 8 + 200 * 200 * 72 * 8 + 100 * 100 * 72 * 8 + 50 * 50 * 72 * 8 + 5008 * 100 + 100 * 50 + 50 * 2
            + 100 * 625 + 50 * 50 * 9 * 2 + 100 * 100 * 18 * 4 + 200 * 200 * 36 * 8 + 400 * 400 * 72)
    assert abs(macs - 77.65e6) < 0.05e6


def test_bilinear_is_tf2_half_pixel_with_edge_clamp():
    x = torch.arange(5, dtype=torch.float32).reshape(1, 1, 1, 5) ** 2
    y = torch.nn.functional.interpolate(x.expand(1, 1, 2, 5), scale_factor=2, mode="bilinear", align_corners=False)[0, 0, 0]
    v = x[0, 0, 0]
    for i in range(5):
        lo, hi = v[max(i - 1, 0)], v[min(i + 1, 4)]
        assert torch.isclose(y[2 * i], 0.25 * lo + 0.75 * v[i])
        assert torch.isclose(y[2 * i + 1], 0.75 * v[i] + 0.25 * hi)


def test_decode_is_f_order_unravel_col_row():
    ptr = torch.zeros(1, 400, 400)
    ptr[0, 37, 250] = 5.0                                   # row 37, col 250
    act = torch.tensor([[0.1, 0.7]])
    ia, xy = po.decode(act, ptr)
    assert int(ia) == 1 and xy.tolist() == [[250, 37]]
    assert np.unravel_index(int(np.argmax(ptr[0].numpy())), (400, 400), order="F") == (250, 37)
    # ties -> first flat index
    ptr[0, 10, 3] = 5.0
    assert po.decode(act, ptr)[1].tolist() == [[3, 10]]


def test_flatten_is_nhwc_and_vector_comes_first():
    w = po.init_weights(1)
    for k in w:
        if k.endswith("kernel"):
            w[k] = torch.zeros_like(w[k])
    # make the trunk an identity-ish probe: conv4 passes channel c of its centre tap through
    for i in range(1, 5):
        k = torch.zeros(3, 3, 2 if i == 1 else 8, 8)
        for c in range(2 if i == 1 else 8):
            k[1, 1, c, c] = 1.0
        w["conv%d/kernel" % i] = k
        w["norm%d/var" % i] = torch.ones(8) - po.BN_EPS    # BN scale exactly 1
    img = torch.zeros(1, 400, 400, 2)
    img[0, 16 * 3:16 * 4, 16 * 7:16 * 8, 1] = 1.0           # a 16x16 block of channel 1 -> pooled cell (3, 7)
    vec = torch.arange(8, dtype=torch.float32).reshape(1, 8)
    d1 = torch.zeros(5008, 100)
    d1[3, 0] = 1.0                                          # picks vector[3]
    d1[8 + (3 * 25 + 7) * 8 + 1, 1] = 1.0                   # picks flat[(h=3, w=7), c=1]
    w["dense1/kernel"] = d1
    _, _, inter = po.forward(w, img, vec, return_intermediates=True)
    assert inter["dense1"][0, 0] == 3.0 and inter["dense1"][0, 1] == 1.0


def test_random_bn_weights_give_finite_outputs():
    w = po.init_weights(3, randomize_bn=True)
    img = (torch.rand(2, 400, 400, 2) < 0.01).float()
    act, ptr = po.forward(w, img, torch.rand(2, 8) * 400)
    assert act.shape == (2, 2) and ptr.shape == (2, 400, 400) and torch.isfinite(ptr).all()
