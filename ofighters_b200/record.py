"""Record / replay of batched games: the B200 counterpart of ``OfighterRecord`` (lib/record.py:7-66).

The reference keeps ``initial_state`` (an Observation), the number of agents and one list of actions
per frame, pickles itself to ``*.orec`` and replays by feeding the lists back to ``game.frame``
(``saveFrame / nextFrame / rewind / save / load``; its Battleground side is unfinished upstream --
lib/battleground.py:100-102 calls a non-existent ``Observation.loadBattleground``).  Here the same
interface holds a deterministic ACTION TAPE for N arenas:

    initial state   the full arena state of ``BatchedBattleground.state()`` (ships, lasers, counters)
    actions         int16 [T, N, S, 4] = (shoot, thrust, pointing_x, pointing_y) per frame, pinned host memory
    resets          {frame index: spawn int32 [N, S, 2]} for the MAX_TIME restarts inside the tape

``nextFrame`` replays one frame through the host-buffer entry point (ofb_step_host_async), so a
tape recorded from device bots, from the policy, or from host bots replays bit-identically; the
same file drives the CPU oracle in the tests.  The on-disk ``.orec`` is a numpy ``.npz`` archive
(no pickled objects).
"""
import numpy as np
import torch

from .battleground import BatchedBattleground
from .config import ArenaConfig


class BatchedRecord:
    def __init__(self, battleground, capacity=256):
        """Start recording ``battleground`` from its current state (lib/ofighters.py:238,403)."""
        bg = battleground
        self.n_arenas, self.nb_agents = bg.n_arenas, bg.ships_number
        self.config = bg.config
        self.initial_state = {k: v.cpu().numpy() for k, v in bg.state().items()}
        self.initial_time = bg.time
        self._tape = torch.empty((capacity, self.n_arenas, self.nb_agents, 4), dtype=torch.int16).pin_memory()
        self.n_frames = 0
        self.resets = {}
        self.reading_head = 0
        self.game = None

    # ------------------------------------------------------------------ recording
    @property
    def actions(self):
        """int16 [T, N, S, 4] (host view of the frames recorded so far)."""
        return self._tape[:self.n_frames]

    def saveFrame(self, actions):
        """Append one frame's actions (device or host tensor int16 [N,S,4]); call it right before the frame is stepped."""
        if self.n_frames == self._tape.shape[0]:
            grown = torch.empty((2 * self._tape.shape[0],) + tuple(self._tape.shape[1:]), dtype=torch.int16).pin_memory()
            grown[:self.n_frames].copy_(self._tape[:self.n_frames])
            self._tape = grown
        self._tape[self.n_frames].copy_(torch.as_tensor(actions).reshape(self._tape.shape[1:]))
        self.n_frames += 1

    def saveReset(self, spawn_xy):
        """Note a ``Battleground.restart`` (with its spawn draws) before the next recorded frame."""
        self.resets[self.n_frames] = torch.as_tensor(spawn_xy).to("cpu", torch.int32).numpy().copy()

    # ------------------------------------------------------------------ replay
    def rewind(self, device=None):
        """Rebuild the game at the initial state and reset the reading head (lib/record.py:38-40)."""
        self.reading_head = 0
        S = self.nb_agents
        bg = BatchedBattleground(self.n_arenas, ships={"external": S}, config=self.config, device=device,
                                 spawn_xy=torch.zeros((self.n_arenas, S, 2), dtype=torch.int32))
        bg.load_state({k: torch.from_numpy(v) for k, v in self.initial_state.items() if k != "ship_steps"})
        bg.time = self.initial_time
        _refresh_obs(bg)
        self.game = bg
        self._obs = [torch.empty((self.n_arenas, S, 8), dtype=torch.float32).pin_memory() for _ in range(2)]
        return bg

    def nextFrame(self):
        """Replay one recorded frame; returns the game (its ``obs_host`` holds the observation heads after
        ``game.wait_host()``)."""
        if self.game is None:
            self.rewind()
        k = self.reading_head
        if k >= self.n_frames:
            raise Exception("end of record")
        if k in self.resets:
            self.game.restart(spawn_xy=torch.from_numpy(self.resets[k]))
        self.obs_host = self._obs[k & 1]
        self.game.step_host(self._tape[k], self.obs_host, wait=False)
        self.reading_head += 1
        return self.game

    def __str__(self):
        return "BatchedRecord(%d frames x %d arenas x %d agents)" % (self.n_frames, self.n_arenas, self.nb_agents)

    # ------------------------------------------------------------------ file format
    def save(self, name):
        """Save the record without the game object (lib/record.py:47-56)."""
        if not name.endswith(".orec"):
            name += ".orec"
        arrays = {"init/" + k: v for k, v in self.initial_state.items()}
        arrays["actions"] = self.actions.numpy()
        arrays["meta"] = np.array([self.n_arenas, self.nb_agents, self.initial_time, self.n_frames], dtype=np.int64)
        arrays["config"] = np.array([getattr(self.config, f) for f in _CONFIG_FIELDS], dtype=np.int64)
        for k, sp in self.resets.items():
            arrays["reset/%d" % k] = sp
        with open(name, "wb") as f:
            np.savez_compressed(f, **arrays)
        return name

    @classmethod
    def load(cls, name):
        """Load a record; ``rewind()`` / ``nextFrame()`` start the replay (lib/record.py:59-66)."""
        if not name.endswith(".orec"):
            name += ".orec"
        with np.load(name) as z:
            rec = cls.__new__(cls)
            rec.n_arenas, rec.nb_agents, rec.initial_time, rec.n_frames = (int(v) for v in z["meta"])
            rec.config = ArenaConfig(**{f: int(v) for f, v in zip(_CONFIG_FIELDS, z["config"])})
            rec.initial_state = {k[5:]: z[k] for k in z.files if k.startswith("init/")}
            rec.resets = {int(k[6:]): z[k] for k in z.files if k.startswith("reset/")}
            acts = torch.from_numpy(z["actions"].astype(np.int16))
        rec._tape = torch.empty((max(1, rec.n_frames),) + tuple(acts.shape[1:]), dtype=torch.int16).pin_memory() \
            if torch.cuda.is_available() else torch.empty((max(1, rec.n_frames),) + tuple(acts.shape[1:]), dtype=torch.int16)
        rec._tape[:rec.n_frames].copy_(acts)
        rec.reading_head, rec.game = 0, None
        return rec


_CONFIG_FIELDS = ("n_ships", "laser_cap", "width", "height", "max_time", "reward_kill", "reward_death", "reward_aim",
                  "reward_trajectory")


def _refresh_obs(bg):
    import ctypes as C
    from . import _lib
    _lib.check(bg._lib.ofb_obs_vec(bg._h, C.c_void_p(bg.obs_vec.data_ptr()), bg._stream()))
