"""torch-CPU fp32 restatement of the Q-learning update of the bi-head pointer model -- TEST INFRASTRUCTURE, not product.

Follows /root/reference/ofighters/agents/qlearnIA_V2.py: ``Trainer.replay`` (:240-287) builds the TD targets from
``model.predict`` on obs / next_obs and calls ``model.fit(x, y, epochs=1, batch_size=B)`` = ONE optimiser step of the model
compiled with ``loss='mse', optimizer=Adam(lr)`` (:188) -- Keras sums the two outputs' mean-squared errors, runs
BatchNormalization in training mode (batch statistics; moving averages updated with momentum 0.99) and applies Adam.

PARITY UNPINNED: Keras / TensorFlow are un-pinned third-party dependencies that are absent here and the reference's tests hold
no expected values for the update, so this file is pinned only by the published algorithms it restates:
  * Keras 2.x ``Adam.get_updates``: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); m, v exponential averages;
    p -= lr_t * m / (sqrt(v) + epsilon), epsilon = K.epsilon() = 1e-7 (NOT torch.optim.Adam's epsilon placement);
  * ``BatchNormalization`` (non-fused): normalise with the batch mean and the BIASED batch variance, epsilon 1e-3;
    moving = moving * momentum + batch * (1 - momentum)  (``unbiased_moving_var`` switches to TF's fused-kernel n/(n-1));
  * ``mse`` = mean over every element of the output; total loss = sum over the two outputs.
Gradients come from torch autograd on the same forward as oracle/policy_torch.py written with batch statistics.
"""
import torch
import torch.nn.functional as F

from . import policy_torch as po

BN_MOMENTUM = 0.99
ADAM = dict(lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-7)
GAMMA = 0.9                                              # Trainer.gamma, qlearnIA_V2.py:51

BN_LAYERS = ["norm1", "norm2", "norm3", "norm4", "upnorm1", "upnorm2", "upnorm3"]


def trainable_names(w):
    return [k for k in w if not (k.endswith("/mean") or k.endswith("/var"))]


def _conv_bn_relu_train(x, w, conv, bn, stats):
    k = w[conv + "/kernel"].permute(3, 2, 0, 1).contiguous()          # HWIO -> OIHW
    y = F.conv2d(x, k, w[conv + "/bias"], padding=1)
    if bn is not None:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        stats[bn] = (mean.detach(), var.detach(), y.numel() // y.shape[1])
        xhat = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + po.BN_EPS)
        y = F.relu(xhat * w[bn + "/gamma"][None, :, None, None] + w[bn + "/beta"][None, :, None, None])
    return y


def forward_train(w, image, vector):
    """Training-mode forward (batch statistics).  image [B,400,400,2], vector [B,8] -> act [B,2], ptr [B,400,400], stats."""
    stats = {}
    dt = w["conv1/kernel"].dtype                         # fp32 (Keras' dtype); the tests' finite differences use fp64
    x = image.to(dt).permute(0, 3, 1, 2).contiguous()
    for i in range(1, 5):
        x = _conv_bn_relu_train(x, w, "conv%d" % i, "norm%d" % i, stats)
        x = F.max_pool2d(x, 2)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    cat = torch.cat([vector.to(dt), flat], dim=1)
    h = F.relu(cat @ w["dense1/kernel"] + w["dense1/bias"])
    d2 = F.relu(h @ w["dense2/kernel"] + w["dense2/bias"])
    act = d2 @ w["output1/kernel"] + w["output1/bias"]
    u = F.relu(h @ w["updense1/kernel"] + w["updense1/bias"])
    u = u.reshape(-1, 25, 25, 1).permute(0, 3, 1, 2)
    for i in range(1, 5):
        u = F.interpolate(u, scale_factor=2, mode="bilinear", align_corners=False)
        u = _conv_bn_relu_train(u, w, "upconv%d" % i, "upnorm%d" % i if i < 4 else None, stats)
    return act, u[:, 0], stats


def loss_and_grads(w, image, vector, target_act, target_ptr):
    """-> (total, mse_act, mse_ptr) floats, {name: grad} for the trainable tensors, BN batch stats."""
    wt = {k: v.clone().detach().requires_grad_(k in trainable_names(w)) for k, v in w.items()}
    act, ptr, stats = forward_train(wt, image, vector)
    la = ((act - target_act) ** 2).mean()
    lp = ((ptr - target_ptr) ** 2).mean()
    (la + lp).backward()
    grads = {k: (wt[k].grad.detach() if wt[k].grad is not None else torch.zeros_like(wt[k])) for k in trainable_names(w)}
    return (float((la + lp).detach()), float(la.detach()), float(lp.detach())), grads, stats


class KerasAdam:
    """Keras 2.x Adam.get_updates restated tensor by tensor."""

    def __init__(self, lr=ADAM["lr"], beta1=ADAM["beta1"], beta2=ADAM["beta2"], eps=ADAM["eps"]):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, w, grads):
        self.t += 1
        lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            m = self.m.get(k, torch.zeros_like(g)) * self.b1 + (1.0 - self.b1) * g
            v = self.v.get(k, torch.zeros_like(g)) * self.b2 + (1.0 - self.b2) * g * g
            self.m[k], self.v[k] = m, v
            w[k] = w[k] - lr_t * m / (torch.sqrt(v) + self.eps)


def fit(w, opt, image, vector, target_act, target_ptr, unbiased_moving_var=False):
    """One ``model.fit`` step in place on ``w``; returns the losses history['loss'] would hold (pre-update)."""
    losses, grads, stats = loss_and_grads(w, image, vector, target_act, target_ptr)
    for bn, (mean, var, n) in stats.items():
        if unbiased_moving_var and n > 1:
            var = var * n / (n - 1)
        w[bn + "/mean"] = w[bn + "/mean"] * BN_MOMENTUM + mean * (1.0 - BN_MOMENTUM)
        w[bn + "/var"] = w[bn + "/var"] * BN_MOMENTUM + var * (1.0 - BN_MOMENTUM)
    opt.step(w, grads)
    return losses


def td_targets(act_obs, ptr_obs, act_next, ptr_next, iaction, pointer, reward, done, gamma=GAMMA):
    """``Trainer.replay`` :262-280.  pointer[b] = (x, y); the reference indexes the [row, col] map with that tuple,
    i.e. writes ptr_target[x][y] -- reproduced."""
    t_act, t_ptr = act_obs.clone(), ptr_obs.clone()
    g = torch.tensor(gamma, dtype=torch.float32)         # fp32 arithmetic like Keras' arrays: (gamma * max) * keep + reward
    for b in range(act_obs.shape[0]):
        keep = torch.tensor(0.0 if bool(done[b]) else 1.0)
        r = reward[b].to(torch.float32)
        t_act[b, int(iaction[b])] = r + g * act_next[b].max() * keep
        t_ptr[b, int(pointer[b, 0]), int(pointer[b, 1])] = r + g * ptr_next[b].max() * keep
    return t_act, t_ptr


def flatten_weights(w, spec):
    return torch.cat([w[name].reshape(-1).to(torch.float32) for name, _ in spec])


def unflatten_weights(flat, spec):
    out, off = {}, 0
    for name, shape in spec:
        n = 1
        for d in shape:
            n *= d
        out[name] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out
