"""B200 counterpart of the reference's ``Trainer`` update path (agents/qlearnIA_V2.py:47-299) and its epsilon
schedules (lib/epsilon.py:18-86) -- SURVEY.md 8(f) rank 1.

    reference                                             here
    ----------------------------------------------------  ---------------------------------------------------
    Trainer(learning_rate, epsilon, batch_size, memory)   TrainerB200(weights=None, learning_rate=..., ...)
    .remember(state, iaction, ipointer, reward, next, d)  .remember(...)   state = (maps_bits [2,5000], vec [8])
    .replay(batch_size) -> History (loss)                 .replay(batch_size) -> History-like (.history["loss"])
    .get_best_action(obs, rand=True)                      .get_best_action(maps_bits, vec, rand=True)
    .decay_epsilon()                                      .decay_epsilon()
    model.fit(x, y, epochs=1, batch_size=B)               .fit(maps_bits, vec, target_act, target_ptr)
    Epsilon_cos(period) / Epsilon_decay()                 epsilon.EpsilonSchedule (closed form in t, evaluated on the device)
    .save(id)  /  load_model(name)                        .save(id) -> .npz of Keras-layout arrays / TrainerB200.load(path)

``predict`` runs on the inference engine (PolicyB200: tcgen05 kernels, BN folded, moving statistics); ``fit`` runs
libofb's training kernels (fp32, BatchNormalization on batch statistics, Keras' Adam) through include/ofb_train.h and
then refreshes the inference engine's weights.  There is no torch / CPU fallback for either.
"""
import ctypes as C
import os
import random
import time
from collections import deque
from types import SimpleNamespace

import torch

from . import _lib
from .epsilon import Epsilon_cos, Epsilon_decay, EpsilonSchedule  # noqa: F401  (re-exported: Trainer(epsilon=Epsilon_cos(...)))
from .policy import WEIGHT_SPEC, PolicyB200, keras_default_weights


def flatten_weights(weights):
    """Keras-layout dict -> the flat fp32 array of include/ofb_train.h (WEIGHT_SPEC order)."""
    parts = []
    for name, shape in WEIGHT_SPEC:
        if name not in weights:
            raise Exception("missing weight " + name)
        t = torch.as_tensor(weights[name]).detach().to("cpu", torch.float32).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise Exception("weight {} : expected shape {} but got shape {}.".format(name, shape, tuple(t.shape)))
        parts.append(t.reshape(-1))
    return torch.cat(parts).contiguous()


def unflatten_weights(flat):
    out, off = {}, 0
    for name, shape in WEIGHT_SPEC:
        n = 1
        for d in shape:
            n *= d
        out[name] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class TrainerB200:
    def __init__(self, weights=None, name=None, learning_rate=0.001, epsilon=None, batch_size=30, memory_size=400,
                 device=None, seed=0, max_ships=64, bn_unbiased_moving_var=False):
        """Defaults are the reference constructor's (qlearnIA_V2.py:48); the module-level TRAINER is built with
        ``learning_rate=0.0001, epsilon=Epsilon_cos(period=110*400), batch_size=8`` (:304-310)."""
        if not torch.cuda.is_available():
            raise _lib.OfbError("TrainerB200 needs a CUDA device: fit and predict are hand-written sm_100a kernels "
                                "and have no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._lib = _lib.load()
        self.state_size = 320008
        self.action_size = 2
        self.gamma = 0.9                                 # :51
        self.epsilon = epsilon if epsilon is not None else Epsilon_decay()
        self.learning_rate = learning_rate
        self.memory = deque(maxlen=memory_size)
        self.batch_size = batch_size
        self.name = name
        self.ptr_values = None
        self.act_values = None
        self.launch_count = 0
        weights = weights if weights is not None else keras_default_weights(seed)
        flat = flatten_weights(weights)
        cfg = _lib.OfbTrainConfig()
        self._lib.ofb_train_default_config(C.byref(cfg))
        cfg.lr = float(learning_rate)
        cfg.max_batch = max(int(batch_size), 1)
        cfg.bn_unbiased_moving_var = 1 if bn_unbiased_moving_var else 0
        h = C.c_void_p()
        _lib.check(self._lib.ofb_trainer_create(C.c_void_p(flat.data_ptr()), flat.numel(), C.byref(cfg),
                                                self.device.index or 0, C.byref(h)))
        self._h = h
        self.max_batch = cfg.max_batch
        self.model = PolicyB200(weights, device=self.device, max_ships=max_ships)   # the predict() side
        self._loss = torch.zeros(3, dtype=torch.float32, device=self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.ofb_trainer_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ weights
    def get_weights(self):
        flat = torch.empty(len_flat(), dtype=torch.float32)
        _lib.check(self._lib.ofb_trainer_get_weights(self._h, C.c_void_p(flat.data_ptr()), self._stream()))
        return unflatten_weights(flat)

    def get_grads(self):
        flat = torch.empty(len_flat(), dtype=torch.float32)
        _lib.check(self._lib.ofb_trainer_get_grads(self._h, C.c_void_p(flat.data_ptr()), self._stream()))
        return unflatten_weights(flat)

    @property
    def steps(self):
        return int(self._lib.ofb_trainer_steps(self._h))

    def sync_model(self):
        """Hand the trained weights to the inference engine (BN folded from the moving statistics)."""
        self.model.load_weights(self.get_weights())

    def save(self, id=None, overwrite=False, folder="networks"):
        """``Trainer.save`` (:289-298): same naming -- ``keras-model[-<name> | -<timestamp>][-<id>]`` under ``folder`` -- but
        the file is an ``.npz`` of the Keras-layout arrays of WEIGHT_SPEC (what ``PolicyB200.load_weights`` / ``load`` take);
        Keras' HDF5 container is out of scope.  Returns the path."""
        import numpy as np
        name = "keras-model-" + (self.name if self.name else time.strftime("%Y-%m-%d_%H-%M-%S"))
        if id:
            name += "-" + str(id)
        os.makedirs(folder, exist_ok=True)
        path = os.path.join(folder, name + ".npz")
        if os.path.exists(path) and not overwrite:
            raise Exception("{} exists and overwrite is False.".format(path))
        np.savez(path, **{k.replace("/", "."): v.numpy() for k, v in self.get_weights().items()})
        return path

    @staticmethod
    def load_weights_file(path):
        """Inverse of ``save``: the Keras-layout dict (``load_model`` of :70 for this format)."""
        import numpy as np
        with np.load(path) as f:
            return {k.replace(".", "/"): torch.from_numpy(f[k]) for k in f.files}

    # ------------------------------------------------------------------ fit / predict
    def _check_batch(self, maps_bits, vec):
        B = maps_bits.shape[0]
        if maps_bits.dtype != torch.int32 or tuple(maps_bits.shape) != (B, 2, 5000) or not maps_bits.is_contiguous() \
                or maps_bits.device != self.device:
            raise Exception("Invalid maps : expected contiguous int32 [B, 2, 5000] bit maps on {}.".format(self.device))
        if B < 1 or B > self.max_batch:
            raise Exception("Invalid batch : {} samples but the trainer was sized for {}.".format(B, self.max_batch))
        vec = vec.to(device=self.device, dtype=torch.float32).contiguous().reshape(-1, 8)
        if vec.shape[0] != B:
            raise Exception("Invalid vector input : expected shape {} but got shape {}.".format((B, 8), tuple(vec.shape)))
        return B, vec

    def fit(self, maps_bits, vec, target_act, target_ptr, sync_model=True):
        """``model.fit(x=[img, vec], y=[act, ptr], epochs=1, batch_size=B)`` (:286): one Adam step.  Returns the loss
        tensor (total, mse(output1), mse(output2)) of the batch before the update (device, float32 [3])."""
        B, vec = self._check_batch(maps_bits, vec)
        target_act = target_act.to(device=self.device, dtype=torch.float32).contiguous()
        target_ptr = target_ptr.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(target_act.shape) != (B, 2) or target_ptr.numel() != B * 160000:
            raise Exception("Invalid targets : expected shapes {} and {}.".format((B, 2), (B, 400, 400, 1)))
        _lib.check(self._lib.ofb_trainer_fit(self._h, _ptr(maps_bits), _ptr(vec), _ptr(target_act), _ptr(target_ptr), B,
                                             _ptr(self._loss), self._stream()))
        self.launch_count += self.KERNELS_PER_FIT
        loss = self._loss.clone()
        if sync_model:
            self.sync_model()
        return loss

    KERNELS_PER_FIT = 96          # forward 33 + backward 62 + Adam (memsets not counted)

    def forward_train(self, maps_bits, vec):
        """Training-mode predictions (batch statistics) -- what ``fit`` differentiates.  -> act [B,2], ptr [B,400,400]."""
        B, vec = self._check_batch(maps_bits, vec)
        act = torch.empty((B, 2), dtype=torch.float32, device=self.device)
        ptr = torch.empty((B, 400, 400), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.ofb_trainer_forward(self._h, _ptr(maps_bits), _ptr(vec), B, _ptr(act), _ptr(ptr), self._stream()))
        return act, ptr

    def predict(self, maps_bits, vec):
        """``model.predict`` on bit maps: (act [B,2], ptr [B,400,400]) from the inference engine."""
        r = self.model.forward(maps_bits, vec, 1, want_ptr=True)
        return r["act"], r["ptr"]

    # ------------------------------------------------------------------ reference surface
    def decay_epsilon(self):
        self.epsilon.next()

    def random_play(self):
        """qlearnIA_V2.py:317-321"""
        iaction = random.randint(0, self.action_size - 1)
        return [iaction, (random.randint(0, 400 - 1), random.randint(0, 400 - 1))]

    def get_best_action(self, maps_bits, vec, rand=True):
        """:199-235 for one observation: maps_bits [2,5000] (or [1,2,5000]), vec [8] -> [iaction, (x, y)]."""
        if rand and random.random() <= self.epsilon.get():
            return self.random_play()
        r = self.model.forward(maps_bits.reshape(1, 2, 5000).contiguous(), vec.reshape(1, 8), 1, want_ptr=True)
        self.act_values = r["act"][0]
        self.ptr_values = r["ptr"][0]
        xy = r["xy"][0].tolist()
        return [int(r["iaction"][0]), (int(xy[0]), int(xy[1]))]

    def remember(self, state, iaction, ipointer, reward, next_state, done):
        """:237-238.  ``state`` / ``next_state`` = (maps_bits int32 [2,5000], vec float32 [8]) device tensors."""
        self.memory.append([state, iaction, ipointer, reward, next_state, done])

    def replay(self, batch_size):
        """:240-287.  Targets from predictions on obs and next_obs; like the reference, the fitted INPUTS are the
        next observations (``inputs1[i] = img_input`` is evaluated after img_input was rebuilt from next_obs, :281-282)."""
        batch_size = min(batch_size, len(self.memory))
        if batch_size < 1:
            raise Exception("replay on an empty memory")
        minibatch = random.sample(self.memory, batch_size)
        dev = self.device
        maps_o = torch.stack([m[0][0] for m in minibatch]).to(dev).contiguous()
        vec_o = torch.stack([m[0][1] for m in minibatch]).to(dev).contiguous()
        maps_n = torch.stack([m[4][0] for m in minibatch]).to(dev).contiguous()
        vec_n = torch.stack([m[4][1] for m in minibatch]).to(dev).contiguous()
        iaction = torch.tensor([int(m[1]) for m in minibatch], dtype=torch.int32, device=dev)
        pointer = torch.tensor([[int(m[2][0]), int(m[2][1])] for m in minibatch], dtype=torch.int32, device=dev)
        reward = torch.tensor([float(m[3]) for m in minibatch], dtype=torch.float32, device=dev)
        done = torch.tensor([1 if m[5] else 0 for m in minibatch], dtype=torch.uint8, device=dev)
        act_o, ptr_o = self.predict(maps_o, vec_o)
        act_n, ptr_n = self.predict(maps_n, vec_n)
        _lib.check(self._lib.ofb_trainer_td_targets(_ptr(act_o), _ptr(ptr_o), _ptr(act_n), _ptr(ptr_n), _ptr(iaction),
                                                    _ptr(pointer), _ptr(reward), _ptr(done), float(self.gamma), batch_size,
                                                    _ptr(act_o), _ptr(ptr_o), self._stream()))
        self.launch_count += 1
        loss = self.fit(maps_n, vec_n, act_o, ptr_o)
        host = loss.cpu().tolist()
        return SimpleNamespace(history={"loss": [host[0]], "output1_loss": [host[1]], "output2_loss": [host[2]]})


def len_flat():
    n = 0
    for _, shape in WEIGHT_SPEC:
        k = 1
        for d in shape:
            k *= d
        n += k
    return n


class QLearner:
    """Batched counterpart of ``QlearnIA.play``'s learning bookkeeping (agents/qlearnIA_V2.py:372-417) on top of one shared
    ``TrainerB200`` ("All bots share the same trainer", :362).  The policy ship of each of the first ``track`` arenas is a
    QlearnIA bot with ids 1..track (bot id 1 = arena 0): every frame it remembers (previous_obs, previous_action,
    previous_pointer, obs.reward, obs, obs.done) until it has seen its own death (:376-392); ANY bot replays when it dies
    (:378-385); bot id 1 alone advances epsilon -- once per decision it takes, i.e. not after its death and not during the
    collecting phase (:394-401) -- replays every ``replay_every`` total steps (:403-405) and snapshots the model every
    ``snapshot`` episodes (:416-417).  During the first ``collecting_steps`` total steps every policy ship plays
    ``random_play()`` (:394-396).  Observations stay on the device: obs = (maps_bits [2,5000], head [8]).

    ``actor``: the inference engine that drives the arenas (default: the trainer's own ``model``).  With a separate actor
    (multi-GPU: every rank's actor is refreshed from the broadcast weights at common points) the trainer's own small
    ``model`` keeps serving ``replay``'s predictions with the latest weights."""

    def __init__(self, trainer, track=8, replay_every=50, collecting_steps=20, snapshot=50, actor=None, snapshot_folder="networks"):
        self.trainer = trainer
        self.actor = actor if actor is not None else trainer.model
        self.track = int(track)
        self.replay_every = int(replay_every)
        self.collecting_steps = int(collecting_steps)
        self.snapshot = int(snapshot)
        self.snapshot_folder = snapshot_folder
        self.total_steps = 0          # Agent.total_steps of bot 1 (never reset, agents/agent.py:66-68)
        self.steps = 0                # Agent.steps: frames since the last reset
        self.episode = 0
        self.losses = []
        self.epsilons = []
        self.saved = []
        self.previous = [None] * self.track
        self.done = [False] * self.track

    @property
    def collecting(self):
        """True while the NEXT decision belongs to the random collecting phase (``total_steps < collecting_steps`` is
        evaluated after Agent.step incremented the counter)."""
        return self.total_steps + 1 < self.collecting_steps

    def reset(self):
        """``QlearnIA.reset`` (:359-369) + ``Agent.reset`` (agents/agent.py:59-64): log epsilon, forget the previous action."""
        self.episode += 1
        self.steps = 0
        self.epsilons.append(self.trainer.epsilon.get())
        self.previous = [None] * self.track
        self.done = [False] * self.track

    def act(self, bg, maps_bits):
        """One frame of every QlearnIA ship: choose (collecting phase / eps-greedy with the trainer's schedule on the
        device), write the action rows, then the bookkeeping of ``observe`` with the actions that are PLAYED.
        -> (iaction, xy, replayed)"""
        collecting = self.collecting
        iact, xy = self.actor.act(bg, maps_bits, epsilon=self.trainer.epsilon, collecting=collecting)
        return iact, xy, self.observe(bg, maps_bits, iact, xy)

    def observe(self, bg, maps_bits, iaction, xy, ship=None):
        """Call once per frame right after the policy ships' PLAYED actions (iaction [A*P], xy [A*P,2]) were written for
        the current observation.  Returns True when the trainer's weights changed (a replay ran)."""
        if ship is None:
            ship = int(bg._policy_ship_idx[0]) if getattr(bg, "_policy_ship_idx", None) is not None else 0
        K = min(self.track, bg.n_arenas)
        P = iaction.numel() // bg.n_arenas
        heads = bg.obs_vec[:K, ship, :]
        alive = bg.state(("ship_alive",))["ship_alive"][:K, ship].tolist()
        rewards = heads[:, 0].tolist()
        ia = iaction.reshape(bg.n_arenas, P)[:K, 0].tolist()
        pt = xy.reshape(bg.n_arenas, P, 2)[:K, 0].tolist()
        self.total_steps += 1
        self.steps += 1
        collecting = self.total_steps < self.collecting_steps
        replayed = False
        for k in range(K):
            if self.done[k]:                              # `if self.done: return None` (:375)
                continue
            obs = (maps_bits[k].clone(), heads[k].clone())
            done = not alive[k]
            if done:                                      # every bot replays on its own death (:376-385)
                replayed |= self._replay()
                self.done[k] = True
            if self.previous[k] is not None:
                po, pa, pp = self.previous[k]
                self.trainer.remember(po, pa, pp, rewards[k], obs, done)
            self.previous[k] = (obs, int(ia[k]), (int(pt[k][0]), int(pt[k][1])))
            if k == 0:                                    # "All bots share the same trainer so we only save it once"
                if not collecting:
                    self.trainer.decay_epsilon()          # (:398-401)
                if self.total_steps % self.replay_every == 0:
                    replayed |= self._replay()            # (:403-405)
                if self.snapshot > 0 and self.episode > 0 and self.episode % self.snapshot == 0 and self.steps < 2:
                    self.saved.append(self.trainer.save(id="iteration-%s" % self.episode, overwrite=True,
                                                        folder=self.snapshot_folder))     # (:416-417)
        return replayed

    def _replay(self):
        if len(self.trainer.memory) == 0:
            return False
        h = self.trainer.replay(self.trainer.batch_size)
        self.losses.append(h.history["loss"][0])
        return True
