"""B200 counterpart of the reference's ``Trainer`` forward path (agents/qlearnIA_V2.py:123-235).

    reference                                          here
    -------------------------------------------------  ------------------------------------------------
    Trainer.model (Keras bi-head pointer model)        PolicyB200(weights)   weights = Keras-layout dict
    model.predict([img(B,400,400,2), vec(B,8)])        PolicyB200.predict([img, vec]) -> [act, ptr]
    Trainer.get_best_action(obs) -> [iaction, (x, y)]  PolicyB200.forward_argmax(maps_bits, vec)
    QlearnIA.play -> Action(vector=[s, t, x, y])       PolicyB200.act(bg, maps)  (writes the action rows)

The weights dict uses the names of the Keras layers in definition order (``conv1/kernel`` HWIO,
``norm1/gamma`` ..., ``dense1/kernel`` (in, out) ...; see ``WEIGHT_SPEC``).  All compute runs in
libofb.so (tcgen05 tensor-core kernels + CUDA-core glue); there is no torch / CPU fallback.
"""
import ctypes as C
import math
import os

import torch

from . import _lib

# (name, shape) in Keras layer order
WEIGHT_SPEC = []
for _i, (_cin, _cout) in enumerate([(2, 8), (8, 8), (8, 8), (8, 8)], 1):
    WEIGHT_SPEC += [("conv%d/kernel" % _i, (3, 3, _cin, _cout)), ("conv%d/bias" % _i, (_cout,))]
    WEIGHT_SPEC += [("norm%d/%s" % (_i, k), (_cout,)) for k in ("gamma", "beta", "mean", "var")]
for _n, _fin, _fout in [("dense1", 5008, 100), ("dense2", 100, 50), ("output1", 50, 2), ("updense1", 100, 625)]:
    WEIGHT_SPEC += [(_n + "/kernel", (_fin, _fout)), (_n + "/bias", (_fout,))]
for _i, (_cin, _cout) in enumerate([(1, 2), (2, 4), (4, 8), (8, 1)], 1):
    WEIGHT_SPEC += [("upconv%d/kernel" % _i, (3, 3, _cin, _cout)), ("upconv%d/bias" % _i, (_cout,))]
    if _i < 4:
        WEIGHT_SPEC += [("upnorm%d/%s" % (_i, k), (_cout,)) for k in ("gamma", "beta", "mean", "var")]


def keras_default_weights(seed=0):
    """Fresh-model weights with the Keras defaults the reference relies on (he_uniform convs,
    glorot_uniform dense, zero bias, BN gamma=1 beta=0 mean=0 var=1), ``torch.Generator(seed)``."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for name, shape in WEIGHT_SPEC:
        kind = name.split("/")[1]
        if kind == "kernel" and len(shape) == 4:
            lim = math.sqrt(6.0 / (9 * shape[2]))
            w[name] = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif kind == "kernel":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            w[name] = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif kind in ("gamma", "var"):
            w[name] = torch.ones(shape)
        else:
            w[name] = torch.zeros(shape)
    return w


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class PolicyB200:
    def __init__(self, weights, device=None, max_ships=1024, bilinear="tf2", fused_tail=True, dense_trunk=False, cc_sparse_trunk=False, fused_trunk=True, tail_pair=False):
        """``bilinear``: "tf2" (half-pixel centres, the default of every TF2 / Keras >= 2.3 UpSampling2D) or "tf1" (the legacy
        asymmetric kernel of TF1.x / standalone Keras 2.2) -- the reference pins no version (requirements.txt:1-2), so the
        choice is the caller's.  ``fused_tail=False`` / ``dense_trunk=True`` / ``tail_pair=True`` select the alternative kernels (measurement aids)."""
        if bilinear not in ("tf2", "tf1"):
            raise Exception("bilinear must be 'tf2' or 'tf1'")
        self.bilinear = bilinear
        self._flags = (_lib.POLICY_BILINEAR_TF1 if bilinear == "tf1" else 0) | (0 if fused_tail else _lib.POLICY_UNFUSED_TAIL) | \
            (_lib.POLICY_DENSE_TRUNK if dense_trunk else 0) | (_lib.POLICY_CC_SPARSE_TRUNK if cc_sparse_trunk else 0) | (0 if fused_trunk else _lib.POLICY_UNFUSED_TRUNK) | \
            (_lib.POLICY_TAIL_PAIR if tail_pair else 0)
        self.fused_tail = bool(fused_tail) and not os.environ.get("OFB_POLICY_UNFUSED_TAIL", "").strip("0")
        self.fused_trunk = bool(fused_trunk) and not dense_trunk and not os.environ.get("OFB_POLICY_UNFUSED_TRUNK", "").strip("0") \
            and not os.environ.get("OFB_POLICY_DENSE_TRUNK", "").strip("0")
        if not torch.cuda.is_available():
            raise _lib.OfbError("PolicyB200 needs a CUDA device: the forward is made of hand-written "
                                "sm_100a kernels and has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._lib = _lib.load()
        self.max_ships = int(max_ships)
        self._h = None
        self._engine = "tensor"
        self.launch_count = 0         # kernels of libofb launched by this object so far
        self.act_values = None        # Trainer.act_values / ptr_values of the last predict (qlearnIA_V2.py:214-215)
        self.ptr_values = None
        self.load_weights(weights)

    @classmethod
    def random_init(cls, device=None, seed=0, **kw):
        return cls(keras_default_weights(seed), device=device, **kw)

    def load_weights(self, weights):
        """``weights``: dict name -> array-like in Keras layouts (see WEIGHT_SPEC)."""
        host = {}
        for name, shape in WEIGHT_SPEC:
            if name not in weights:
                raise Exception("missing weight " + name)
            t = torch.as_tensor(weights[name]).detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise Exception("weight {} : expected shape {} but got shape {}.".format(name, shape, tuple(t.shape)))
            host[name] = t
        w = _lib.OfbPolicyWeights()

        def conv(dst, cname, bname):
            dst.kernel, dst.bias = host[cname + "/kernel"].data_ptr(), host[cname + "/bias"].data_ptr()
            if bname is not None:
                dst.gamma, dst.beta = host[bname + "/gamma"].data_ptr(), host[bname + "/beta"].data_ptr()
                dst.mean, dst.var = host[bname + "/mean"].data_ptr(), host[bname + "/var"].data_ptr()

        for i in range(4):
            conv(w.conv[i], "conv%d" % (i + 1), "norm%d" % (i + 1))
            conv(w.upconv[i], "upconv%d" % (i + 1), "upnorm%d" % (i + 1) if i < 3 else None)
        for n in ("dense1", "dense2", "output1", "updense1"):
            d = getattr(w, n)
            d.kernel, d.bias = host[n + "/kernel"].data_ptr(), host[n + "/bias"].data_ptr()
        if self._h:                                       # same architecture: refresh the resident weights in place
            _lib.check(self._lib.ofb_policy_set_weights(self._h, C.byref(w), self._stream()))
        else:
            h = C.c_void_p()
            _lib.check(self._lib.ofb_policy_create_opts(C.byref(w), self.device.index or 0, self.max_ships, self._flags, C.byref(h)))
            self._h = h
        self.weights = host

    def set_engine(self, name):
        """"tensor" (tcgen05, default) or "cuda_core" (same arithmetic on CUDA cores; validation twin)."""
        _lib.check(self._lib.ofb_policy_set_engine(self._h, {"tensor": 0, "cuda_core": 1}[name]))
        self._engine = name

    LAYERS = ("trunk12", "conv3", "conv4", "dense1", "heads", "up3", "up4", "argmax", "tail")

    @property
    def kernels_per_chunk(self):
        """trunk12, conv3, conv4, dense1, heads + the fused tail (or up3, up4, argmax)."""
        if self._engine != "tensor":
            return 9
        return (1 if self.fused_trunk else 3) + 2 + (1 if self.fused_tail else 3)

    def set_taps(self, enable):
        """Validation only: the fused tail kernel also writes upconv3's output so that ``debug_tap(6, ...)`` can read it."""
        _lib.check(self._lib.ofb_policy_set_taps(self._h, 1 if enable else 0))

    def profile(self, enable):
        """enable=True: start bracketing every kernel with CUDA events.  enable=False: stop and return
        {layer: milliseconds per forward call} averaged over the calls made in between."""
        if enable:
            self._prof_calls = 0
            _lib.check(self._lib.ofb_policy_profile(self._h, 1, None))
            return None
        ms = (C.c_float * 9)()
        _lib.check(self._lib.ofb_policy_profile(self._h, 0, ms))
        calls = max(1, getattr(self, "_prof_calls", 1))
        return {k: float(ms[i]) / calls for i, k in enumerate(self.LAYERS)}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.ofb_policy_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ forward
    def forward(self, maps_bits, vec, ships_per_arena=1, want_ptr=False, want_act=True, want_argmax=True):
        """maps_bits int32 [A,2,5000] (``BatchedBattleground.raster("bits")``), vec float32 [A*P,8].
        Returns dict(act [A*P,2], ptr [A*P,400,400] | None, iaction int32 [A*P], xy int32 [A*P,2])."""
        A = maps_bits.shape[0]
        n = A * ships_per_arena
        if maps_bits.dtype != torch.int32 or tuple(maps_bits.shape) != (A, 2, 5000) or not maps_bits.is_contiguous():
            raise Exception("Invalid maps : expected contiguous int32 [A, 2, 5000] bit maps.")
        vec = vec.to(device=self.device, dtype=torch.float32).contiguous().reshape(-1, 8)
        if vec.shape[0] != n:
            raise Exception("Invalid vector input : expected shape {} but got shape {}.".format((n, 8), tuple(vec.shape)))
        act = torch.empty((n, 2), dtype=torch.float32, device=self.device) if want_act else None
        ptr = torch.empty((n, 400, 400), dtype=torch.float32, device=self.device) if want_ptr else None
        iact = torch.empty((n,), dtype=torch.int32, device=self.device)
        xy = torch.empty((n, 2), dtype=torch.int32, device=self.device) if want_argmax else None
        _lib.check(self._lib.ofb_policy_forward(self._h, _ptr(maps_bits), _ptr(vec), A, ships_per_arena, _ptr(act),
                                                _ptr(ptr), _ptr(iact), _ptr(xy), self._stream()))
        self._prof_calls = getattr(self, "_prof_calls", 0) + 1
        per_chunk = max(1, self.max_ships // ships_per_arena)
        self.launch_count += self.kernels_per_chunk * ((A + per_chunk - 1) // per_chunk)
        return {"act": act, "ptr": ptr, "iaction": iact, "xy": xy}

    def forward_argmax(self, maps_bits, vec, ships_per_arena=1):
        """Fused decode path: no pointer map is materialised.  -> (iaction int32 [B], xy int32 [B,2])."""
        r = self.forward(maps_bits, vec, ships_per_arena, want_ptr=False, want_act=False)
        return r["iaction"], r["xy"]

    def pack_image(self, img):
        """Keras-style image [B,400,400,2] (float32 / bfloat16 / uint8, NHWC, ch0 ship_map) -> bit maps."""
        img = img.to(self.device).contiguous()
        if tuple(img.shape[1:]) != (400, 400, 2):
            raise Exception("Invalid image input : expected shape (B, 400, 400, 2) but got shape {}.".format(tuple(img.shape)))
        fmt = {torch.float32: 3, torch.bfloat16: _lib.OFB_MAP_BF16, torch.uint8: _lib.OFB_MAP_U8}.get(img.dtype)
        if fmt is None:
            img, fmt = img.to(torch.float32), 3
        maps = torch.empty((img.shape[0], 2, 5000), dtype=torch.int32, device=self.device)
        _lib.check(self._lib.ofb_policy_pack_image(_ptr(img), fmt, img.shape[0], _ptr(maps), self._stream()))
        return maps

    def predict(self, inputs):
        """``model.predict([img, vec]) -> [act (B,2), ptr (B,400,400,1)]`` (qlearnIA_V2.py:210).  The
        images on this path are the binary ship/laser masks; a pixel counts as set iff non-zero."""
        img, vec = inputs
        img = torch.as_tensor(img)
        maps = self.pack_image(img)
        r = self.forward(maps, torch.as_tensor(vec), 1, want_ptr=True)
        self.act_values = r["act"][0]
        self.ptr_values = r["ptr"][0]
        return [r["act"], r["ptr"].unsqueeze(-1)]

    def get_best_action(self, maps_bits, vec):
        """``Trainer.get_best_action`` without the epsilon branch: [iaction, (x, y)] per row."""
        return self.forward_argmax(maps_bits, vec)

    # ------------------------------------------------------------------ env glue
    def act(self, bg, maps_bits, epsilon=0.0, collecting=False):
        """Drive the "QlearnIA"/"external" ships of ``bg`` for the coming frame: forward on the
        current maps + observation heads, then QlearnIA.play's action vector (:447-456) straight into
        ``bg.actions``.  ``epsilon``: a float or an ``EpsilonSchedule`` (evaluated on the device at its current step) for
        the eps-greedy random branch (:199-204); ``collecting``: the random phase of :394-396 (no forward at all, like
        the reference).  Returns the (iaction, xy) that are PLAYED (what QlearnIA.play remembers, :399-401)."""
        if collecting:
            n = bg.n_arenas * self._ship_idx(bg).numel()
            iact = torch.zeros((n,), dtype=torch.int32, device=self.device)
            xy = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
        else:
            iact, xy = self.decide(bg, maps_bits)
        self.write(bg, iact, xy, epsilon, collecting)
        return iact, xy

    def _ship_idx(self, bg):
        idx = getattr(bg, "_policy_ship_idx", None)
        if idx is None:
            ids = [i for i, b in enumerate(bg.behaviors) if b in ("QlearnIA", "external")]
            if not ids:
                raise Exception("no policy-driven ship in this battleground")
            idx = bg._policy_ship_idx = torch.tensor(ids, dtype=torch.int32, device=bg.device)
            bg._policy_ship_idx_long = idx.long()
        return idx

    def decide(self, bg, maps_bits):
        """First half of ``act``: queue the forward (asynchronous) -> (iaction [A*P], xy [A*P,2]) on the device.  A host-side
        bot can compute the other ships' rows while it runs."""
        idx = self._ship_idx(bg)
        P = idx.numel()
        vec = bg.obs_vec.index_select(1, bg._policy_ship_idx_long).reshape(-1, 8)
        return self.forward_argmax(maps_bits, vec, P)

    def write(self, bg, iact, xy, epsilon=0.0, collecting=False):
        """Second half of ``act``: the decoded actions become the policy ships' rows of ``bg.actions``; where the eps-greedy
        branch (or the collecting phase) replaced a prediction by ``random_play()``, ``iact`` / ``xy`` are overwritten with
        the action that is played."""
        idx = self._ship_idx(bg)
        if hasattr(epsilon, "as_struct"):
            sched, t = epsilon.as_struct(), float(epsilon.t)
        else:
            sched, t = _lib.OfbEpsSchedule(), 0.0
            sched.kind, sched.start = _lib.EPS_CONST, float(epsilon)
        _lib.check(self._lib.ofb_policy_play_actions(_ptr(iact), _ptr(xy), bg.n_arenas, idx.numel(), _ptr(idx), bg.ships_number,
                                                     C.byref(sched), t, 1 if collecting else 0, bg.seed, bg.arena0,
                                                     bg.total_steps, _ptr(bg.actions), None, self._stream()))
        self.launch_count += 1

    _TAP_STRIDE = {0: 320000, 1: 80000, 2: 20000, 3: 5120, 4: 100, 5: 80000, 6: 320000}

    def debug_tap(self, which, n_items, shape):
        """Copy of an intermediate activation of the last forward's first chunk (validation only)."""
        stride = self._TAP_STRIDE[which]
        dtype = torch.float32 if which == 4 else torch.bfloat16
        out = torch.empty((n_items, stride), dtype=dtype, device=self.device)
        _lib.check(self._lib.ofb_policy_debug_tap(self._h, which, n_items, _ptr(out), self._stream()))
        n = 1
        for d in shape:
            n *= d
        return out[:, :n].reshape((n_items,) + tuple(shape))
