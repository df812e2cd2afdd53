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


def test_tf1_legacy_bilinear_kernel():
    """TF1.x resize_bilinear(align_corners=False): source = dst / 2 -> out[2i] = in[i], out[2i+1] = (in[i] + in[i+1]) / 2,
    the last odd sample clamps onto in[n-1]."""
    x = torch.arange(5, dtype=torch.float32).reshape(1, 1, 1, 5) ** 2
    y = po.upsample2x(x.expand(1, 1, 2, 5).contiguous(), "tf1")
    v = x[0, 0, 0]
    assert y.shape == (1, 1, 4, 10)
    for i in range(5):
        assert float(y[0, 0, 0, 2 * i]) == float(v[i])
        assert float(y[0, 0, 0, 2 * i + 1]) == float(0.5 * (v[i] + v[min(i + 1, 4)]))
    assert torch.equal(y[0, 0, 0], y[0, 0, 3])            # rows: both input rows are equal


def test_two_independent_restatements_agree():
    """oracle/policy_numpy.py (numpy float64, NHWC, explicit index arithmetic, written from the Keras layer semantics) against
    oracle/policy_torch.py run in float64: every tapped tensor and both outputs within 1e-6 of the tensor's scale, for both
    bilinear kernels and for fresh as well as randomised BatchNormalization statistics; and the fp32 run the GPU tests use
    stays within 2e-5 of them.  Removes single-author risk; parity with Keras itself stays unpinned."""
    from oracle import policy_numpy as pn
    g = torch.Generator().manual_seed(7)
    img = (torch.rand((2, 400, 400, 2), generator=g) < 0.02).float()
    img[0, 100:117, 200:217, 0] = 1.0                       # a blob, like a ship disk
    vec = torch.rand((2, 8), generator=g) * 400
    for seed, rbn in ((0, False), (5, True)):
        w = po.init_weights(seed, randomize_bn=rbn)
        for mode in ("tf2", "tf1"):
            a64, p64, i64 = po.forward(w, img, vec, return_intermediates=True, bilinear=mode, dtype=torch.float64)
            an, pn_, inn = pn.forward({k: v.numpy() for k, v in w.items()}, img.numpy(), vec.numpy(), bilinear=mode,
                                      return_intermediates=True)
            for name in ("pool1", "pool2", "pool3", "pool4", "up1", "up2", "up3", "up4"):
                ref = i64[name].permute(0, 2, 3, 1).numpy()
                assert np.abs(inn[name] - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), (seed, mode, name)
            assert np.abs(an - a64.numpy()).max() <= 1e-6 * max(1.0, np.abs(an).max())
            assert np.abs(pn_ - p64.numpy()).max() <= 1e-6 * max(1.0, np.abs(pn_).max())
            a32, p32 = po.forward(w, img, vec, bilinear=mode)
            assert np.abs(p32.numpy() - pn_).max() <= 2e-5 * max(1.0, np.abs(pn_).max())
            assert np.abs(a32.numpy() - an).max() <= 2e-5 * max(1.0, np.abs(an).max())
            ia, xy = po.decode(a64, p64)
            dec = pn.decode(an, pn_)
            assert [(int(ia[b]), (int(xy[b, 0]), int(xy[b, 1]))) for b in range(2)] == dec
