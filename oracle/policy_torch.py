"""torch-CPU fp32 restatement of the bi-head pointer model -- TEST INFRASTRUCTURE, not product.

Follows /root/reference/ofighters/agents/qlearnIA_V2.py:123-190 layer by layer with the Keras
defaults of SURVEY.md Appendix B, and the decode of ``Trainer.get_best_action`` (:199-235).

PARITY UNPINNED: Keras / TensorFlow are un-pinned third-party dependencies that are absent here,
the reference ships no trained weights and its tests hold no expected outputs for the model, so
this file is pinned only by the layer list itself (and by the hand-checked bilinear / flatten /
argmax conventions in tests/test_policy_oracle.py).  Choices where Keras versions differ:
  * UpSampling2D(interpolation='bilinear'): bilinear="tf2" (default) = TF2 half-pixel centres with edge clamp
    (== F.interpolate(scale_factor=2, mode='bilinear', align_corners=False)); bilinear="tf1" = the TF1.x / standalone
    Keras 2.2 legacy kernel tf.image.resize_bilinear(align_corners=False): source coordinate = dst / 2, i.e.
    out[2i] = in[i], out[2i+1] = (in[i] + in[min(i+1, n-1)]) / 2 per axis;
  * BatchNormalization epsilon = 1e-3, inference mode (moving statistics).

Weights are a flat ``dict name -> tensor`` in Keras layer order with Keras layouts:
conv kernels HWIO ``[3,3,Cin,Cout]``, dense kernels ``[in,out]``, BN ``gamma/beta/mean/var``.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3

# (name, Cin, Cout) -- trunk then pointer head; every conv but the last is followed by BN + ReLU
CONVS = [("conv1", 2, 8), ("conv2", 8, 8), ("conv3", 8, 8), ("conv4", 8, 8),
         ("upconv1", 1, 2), ("upconv2", 2, 4), ("upconv3", 4, 8), ("upconv4", 8, 1)]
DENSES = [("dense1", 5008, 100), ("dense2", 100, 50), ("output1", 50, 2), ("updense1", 100, 625)]


def init_weights(seed=0, randomize_bn=False):
    """Keras-default initialisation (he_uniform convs, glorot_uniform dense, zero bias, fresh BN)
    from ``torch.Generator().manual_seed(seed)``; ``randomize_bn`` draws non-trivial BN statistics
    and biases so that the folding paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def uni(shape, lim):
        return (torch.rand(shape, generator=g) * 2 - 1) * lim

    for name, cin, cout in CONVS:
        w[name + "/kernel"] = uni((3, 3, cin, cout), math.sqrt(6.0 / (9 * cin)))
        w[name + "/bias"] = torch.zeros(cout)
        if name != "upconv4":
            bn = name.replace("upconv", "upnorm").replace("conv", "norm")
            w[bn + "/gamma"] = torch.ones(cout)
            w[bn + "/beta"] = torch.zeros(cout)
            w[bn + "/mean"] = torch.zeros(cout)
            w[bn + "/var"] = torch.ones(cout)
    for name, fin, fout in DENSES:
        w[name + "/kernel"] = uni((fin, fout), math.sqrt(6.0 / (fin + fout)))
        w[name + "/bias"] = torch.zeros(fout)
    if randomize_bn:
        for k in list(w):
            if k.endswith("/bias"):
                w[k] = uni(w[k].shape, 0.1)
            elif k.endswith("/gamma"):
                w[k] = 1.0 + uni(w[k].shape, 0.3)
            elif k.endswith("/beta") or k.endswith("/mean"):
                w[k] = uni(w[k].shape, 0.2)
            elif k.endswith("/var"):
                w[k] = 1.0 + uni(w[k].shape, 0.5)
    return w


def _conv_bn_relu(x, w, conv, bn):
    """x NCHW fp32.  Conv2D 3x3 'same' + bias -> BatchNormalization(inference) -> ReLU."""
    k = w[conv + "/kernel"].permute(3, 2, 0, 1).contiguous()          # HWIO -> OIHW
    y = F.conv2d(x, k, w[conv + "/bias"], padding=1)
    if bn is not None:
        s = w[bn + "/gamma"] / torch.sqrt(w[bn + "/var"] + BN_EPS)
        y = (y - w[bn + "/mean"][None, :, None, None]) * s[None, :, None, None] + w[bn + "/beta"][None, :, None, None]
        y = F.relu(y)
    return y


def upsample2x(u, bilinear="tf2"):
    """UpSampling2D(size=2, interpolation='bilinear') on NCHW (see the module docstring for the two kernels)."""
    if bilinear == "tf2":
        return F.interpolate(u, scale_factor=2, mode="bilinear", align_corners=False)
    if bilinear != "tf1":
        raise ValueError("bilinear must be 'tf2' or 'tf1'")
    for dim in (2, 3):
        n = u.shape[dim]
        nxt = torch.index_select(u, dim, torch.clamp(torch.arange(n) + 1, max=n - 1))
        u = torch.stack([u, 0.5 * (u + nxt)], dim=dim + 1)            # interleave: even outputs = in[i], odd = the mean
        shape = list(u.shape)
        shape[dim:dim + 2] = [2 * n]
        u = u.reshape(shape)
    return u


def forward(w, image, vector, return_intermediates=False, bilinear="tf2", dtype=torch.float32):
    """image [B,400,400,2] (NHWC, ch0 ship_map, ch1 laser_map), vector [B,8]
    -> act [B,2], ptr [B,400,400]   (qlearnIA_V2.py:123-190).  ``dtype=torch.float64`` runs the same graph in double
    precision (cross-check against oracle/policy_numpy.py)."""
    if dtype != torch.float32:
        w = {k: v.to(dtype) for k, v in w.items()}
    x = image.to(dtype).permute(0, 3, 1, 2).contiguous()
    inter = {}
    for i in range(1, 5):                                              # :129-147
        x = _conv_bn_relu(x, w, "conv%d" % i, "norm%d" % i)
        x = F.max_pool2d(x, 2)
        inter["pool%d" % i] = x
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)               # Flatten() on NHWC  :150
    cat = torch.cat([vector.to(dtype), flat], dim=1)                   # [vector, flat]     :154
    h = F.relu(cat @ w["dense1/kernel"] + w["dense1/bias"])            # :155
    inter["dense1"] = h
    d2 = F.relu(h @ w["dense2/kernel"] + w["dense2/bias"])             # :158
    act = d2 @ w["output1/kernel"] + w["output1/bias"]                 # :160
    u = F.relu(h @ w["updense1/kernel"] + w["updense1/bias"])          # :163
    u = u.reshape(-1, 25, 25, 1).permute(0, 3, 1, 2)                   # Reshape((25,25,1)) :164
    inter["updense1"] = u
    for i in range(1, 5):                                              # :166-186
        u = upsample2x(u, bilinear)
        u = _conv_bn_relu(u, w, "upconv%d" % i, "upnorm%d" % i if i < 4 else None)
        inter["up%d" % i] = u
    ptr = u[:, 0]
    if return_intermediates:
        return act, ptr, inter
    return act, ptr


def decode(act, ptr):
    """``Trainer.get_best_action`` (:218-220): iaction = argmax(act); pointer = F-order unravel of
    the flat C-order argmax of ptr[row, col]  ==  (x, y) = (k % 400, k // 400)."""
    iaction = torch.argmax(act, dim=1)
    k = torch.argmax(ptr.reshape(ptr.shape[0], -1), dim=1)
    W = ptr.shape[2]
    return iaction, torch.stack([k % W, k // W], dim=1)
