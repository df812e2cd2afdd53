"""Batched, GPU-resident counterpart of the reference's ``Battleground``.

Mirrors lib/battleground.py:11-173 (same method names, argument meaning and error behaviour)
for N independent arenas stepped in lockstep on one B200:

    reference                                    here (all arenas at once)
    -------------------------------------------  ------------------------------------------
    Battleground(ships=int|dict, largeur, ...)   BatchedBattleground(n_arenas, ships=..., ...)
    .request_actions() -> [Action|None]          .request_actions() -> int16 [N,S,4]
    .generate_frame(actions)                     .generate_frame(actions)
    .frame()                                     .frame()
    .restart()                                   .restart(mask=None, spawn_xy=None)
    .outside(x, y)                               .outside(x, y)
    .ships[i] / .lasers[j] / .absolute_state     .arena(k) -> read-only Battleground-shaped view
    bot.play(obs) -> Action                      bot.play_batch(obs_vec, bg) -> int16 [N,S,4]

All compute happens in libofb.so (hand-written sm_100a kernels) through the C ABI of
include/ofb.h; torch is used only to own device buffers and streams.  There is no CPU path.
"""
import ctypes as C
from types import SimpleNamespace

import torch

from . import _lib
from .config import ArenaConfig

_BEHAVIOR_TO_KIND = {None: "idle", "idle": "idle", "random": "random", "turret": "turret", "runner": "runner",
                     "thrust": "thrust", "shoot": "shoot", "stress": "stress",
                     # ships whose action rows come from elsewhere (policy forward, host bot)
                     "QlearnIA": "external", "external": "external"}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class ScriptedBots:
    """Device-side scripted bots of agents/agent.py:99-155, one kind per ship index."""

    def __init__(self, kinds, seed=0x0F16):
        self.kinds = list(kinds)
        self.seed = int(seed)
        self._kinds_dev = None

    def _ensure_kinds(self, bg):
        if self._kinds_dev is None:
            self._kinds_dev = torch.tensor([_lib.BOT_KINDS[k] for k in self.kinds], dtype=torch.uint8,
                                           device=bg.device)

    def play_batch(self, obs_vec, bg):
        self._ensure_kinds(bg)
        _lib.check(bg._lib.ofb_bot_actions(bg._h, 0, _ptr(self._kinds_dev), self.seed, bg.arena0,
                                           bg.total_steps, _ptr(bg.actions), bg._stream()))
        bg.launch_count += 1
        return bg.actions


class BatchedBattleground:
    def __init__(self, n_arenas, ships=2, largeur=None, hauteur=None, config=None, device=None,
                 spawn_xy=None, seed=0x0F16, arena0=0, bot=None, networks=()):
        """``ships``: (int) that many ships with the default behavior ("random"), or
        (dict) {"behavior": number} like the reference (lib/battleground.py:13-28).
        ``arena0``: global id of this shard's first arena (keys the device RNG so that results do
        not depend on how arenas are sharded over GPUs)."""
        default_behavior = "random"
        if isinstance(ships, dict):
            self.ships_map = dict(ships)
        elif isinstance(ships, int):
            self.ships_map = {default_behavior: ships}
        else:
            raise Exception("ships argument must be int or dict.")
        behaviors = []
        for behavior, number in self.ships_map.items():
            if behavior not in _BEHAVIOR_TO_KIND:
                raise Exception("You must give a bot in parameter or select an existing behavior.")
            behaviors += [behavior] * number
        self.behaviors = behaviors
        self.ships_number = len(behaviors)

        cfg = config or ArenaConfig()
        cfg = ArenaConfig(**{**cfg.__dict__, "n_ships": self.ships_number,
                             "width": largeur or cfg.width, "height": hauteur or cfg.height})
        self.config = cfg
        self.dim = SimpleNamespace(x=cfg.width, y=cfg.height)
        self.n_arenas = int(n_arenas)
        self.arena0 = int(arena0)
        self.seed = int(seed)
        self.networks = list(networks)

        if not torch.cuda.is_available():
            raise _lib.OfbError("BatchedBattleground needs a CUDA device: the arena step is a "
                                "hand-written sm_100a kernel and has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._lib = _lib.load()
        N, S = self.n_arenas, self.ships_number
        self.actions = torch.zeros((N, S, 4), dtype=torch.int16, device=self.device)
        self.obs_vec = torch.empty((N, S, 8), dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(6, dtype=torch.int64, device=self.device)
        self._spawn = torch.empty((N, S, 2), dtype=torch.int32, device=self.device)

        c = _lib.OfbConfig()
        self._lib.ofb_default_config(C.byref(c))
        c.n_ships, c.laser_cap, c.width, c.height, c.max_time = S, cfg.laser_cap, cfg.width, cfg.height, cfg.max_time
        c.reward_kill, c.reward_death = cfg.reward_kill, cfg.reward_death
        c.reward_aim, c.reward_trajectory = cfg.reward_aim, cfg.reward_trajectory
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            if spawn_xy is None:
                self._draw_spawn(0)
            else:
                self._spawn.copy_(torch.as_tensor(spawn_xy).to(torch.int32).reshape(N, S, 2))
            _lib.check(self._lib.ofb_create(C.byref(c), N, self.device.index or 0, _ptr(self._spawn),
                                            self._stream(), C.byref(h)))
        self._h = h
        self.launch_count = 1      # kernels launched through the C ABI so far (k_init)
        self.laser_cap = self._lib.ofb_laser_cap(self._h)
        self.state_stride = int(self._lib.ofb_state_stride(self._h))   # HBM bytes per arena
        self.time = 0              # frames since the last restart (lib/battleground.py:32)
        self.episode = 0
        self.total_steps = 0       # frames since construction: keys the bots' RNG
        self.bot = bot or ScriptedBots([_BEHAVIOR_TO_KIND[b] for b in behaviors], seed=seed)
        _lib.check(self._lib.ofb_obs_vec(self._h, _ptr(self.obs_vec), self._stream()))
        self.launch_count += 1

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _draw_spawn(self, episode):
        """randint(0, W) x randint(0, H) per ship (lib/battleground.py:79-81,114) on the device RNG."""
        cfg = self.config
        _lib.check(self._lib.ofb_random_spawn(self.n_arenas, self.ships_number, cfg.width, cfg.height, self.seed,
                                              self.arena0, episode, _ptr(self._spawn), self._stream()))
        self.launch_count = getattr(self, "launch_count", 0) + 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.ofb_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ reference surface
    def outside(self, x, y):
        """lib/battleground.py:125-126 (works on scalars or tensors)."""
        return (x < 0) | (y < 0) | (x >= self.dim.x) | (y >= self.dim.y)

    def set_ia(self, network):
        """lib/battleground.py:120-122: every ship shares one network."""
        self.networks = [network]

    def request_actions(self):
        """lib/battleground.py:146-150: every ship (wreckage too) is shown its observation head
        and asked for an action; rows of dead ships are ignored by the step."""
        acts = self.bot.play_batch(self.obs_vec, self)
        if acts is not self.actions:
            acts = torch.as_tensor(acts, device=self.device)
            if acts.shape != self.actions.shape:
                raise Exception("Invalid actions : expected shape {} but got shape {}.".format(
                    tuple(self.actions.shape), tuple(acts.shape)))
            self.actions.copy_(acts)
        return self.actions

    def _maps_arg(self, maps):
        N, W, H = self.n_arenas, self.config.width, self.config.height
        if tuple(maps.shape) != (N, 2, W * H // 32) or maps.dtype != torch.int32 or not maps.is_contiguous() \
                or maps.device != self.device:
            raise Exception("Invalid map buffer : expected {} {}.".format((N, 2, W * H // 32), torch.int32))
        return _ptr(maps)

    def generate_frame(self, actions=None, maps=None):
        """lib/battleground.py:153-160 for all arenas (K1).  Also refreshes ``obs_vec``.  With ``maps`` (the "bits"
        buffer of ``raster``) the step and the new observation maps are one fused launch (ofb_frame)."""
        if actions is None:
            actions = self.actions
        if actions.dtype != torch.int16 or not actions.is_contiguous() or actions.device != self.device:
            actions = torch.as_tensor(actions).to(device=self.device, dtype=torch.int16).contiguous()
        if tuple(actions.shape) != (self.n_arenas, self.ships_number, 4):
            raise Exception("Invalid actions : expected shape {} but got shape {}.".format(
                (self.n_arenas, self.ships_number, 4), tuple(actions.shape)))
        if maps is not None:
            _lib.check(self._lib.ofb_frame(self._h, _ptr(actions), _ptr(self.obs_vec), self._maps_arg(maps), self._stream()))
        else:
            _lib.check(self._lib.ofb_step(self._h, _ptr(actions), _ptr(self.obs_vec), self._stream()))
        self.launch_count += 1
        self.time += 1
        self.total_steps += 1

    def step_host(self, actions_host, obs_host=None, wait=True, maps=None):
        """``generate_frame`` for a host-side bot loop: ``actions_host`` int16 [N,S,4] and
        ``obs_host`` float32 [N,S,8] are (pinned) HOST tensors -- the batched form of one
        ``Battleground.frame()`` with Python bots.  ``wait=True``: the copies ride the same stream as
        the step (ofb_step_host) and the call returns once ``obs_host`` holds the next observation
        heads.  ``wait=False``: pipelined (ofb_step_host_async) -- the call only queues the frame; the
        H2D / D2H copies overlap the kernels of neighbouring frames, ``obs_host`` is valid after
        ``wait_host()``.  Meant for replaying an action tape (lib/record.py) or open-loop bots.  ``maps`` (pipelined
        form only): the frame also writes its observation maps there (fused launch, ofb_frame_host_async)."""
        if actions_host.dtype != torch.int16 or tuple(actions_host.shape) != (self.n_arenas, self.ships_number, 4) \
                or actions_host.is_cuda or not actions_host.is_contiguous():
            raise Exception("Invalid actions : expected a contiguous host int16 tensor of shape {}.".format(
                (self.n_arenas, self.ships_number, 4)))
        compact = obs_host is not None and obs_host.dtype == torch.int16
        want = (self.n_arenas, self.ships_number, 5 if compact else 8)
        if obs_host is not None and (obs_host.dtype not in (torch.float32, torch.int16) or obs_host.is_cuda or tuple(obs_host.shape) != want):
            raise Exception("Invalid observation buffer : expected host float32 {} or int16 {}.".format(
                (self.n_arenas, self.ships_number, 8), (self.n_arenas, self.ships_number, 5)))
        obs_p = C.c_void_p(obs_host.data_ptr()) if obs_host is not None else None
        if compact and (maps is None or wait):
            raise Exception("compact (int16) observation heads are only supported by the pipelined fused form (wait=False, maps=...).")
        if maps is not None:
            if wait:
                raise Exception("maps= is only supported by the pipelined form (wait=False).")
            fn = self._lib.ofb_frame_host_async_i16 if compact else self._lib.ofb_frame_host_async
            _lib.check(fn(self._h, C.c_void_p(actions_host.data_ptr()), obs_p, self._maps_arg(maps), self._stream()))
            self.launch_count += 1 if compact else 0
        else:
            fn = self._lib.ofb_step_host if wait else self._lib.ofb_step_host_async
            _lib.check(fn(self._h, C.c_void_p(actions_host.data_ptr()), obs_p, self._stream()))
        self.launch_count += 1
        self.time += 1
        self.total_steps += 1
        if wait:
            torch.cuda.current_stream(self.device).synchronize()

    OBS_COMPACT_COLUMNS = (0, 2, 3, 6, 7)      # reward, pointing.x, pointing.y, pos.x, pos.y (lib/observation.py:119-123)

    def obs_compact(self, out=None):
        """The observation heads without their constant entries (can_shoot = 1, dim = the map size), as int16 [N,S,5] on the
        device: what a host-side consumer needs per frame at 10 instead of 32 bytes per ship."""
        if out is None:
            out = torch.empty((self.n_arenas, self.ships_number, 5), dtype=torch.int16, device=self.device)
        _lib.check(self._lib.ofb_obs_pack_i16(_ptr(self.obs_vec), _ptr(out), self.n_arenas * self.ships_number, self._stream()))
        self.launch_count += 1
        return out

    def expand_obs(self, obs16):
        """Inverse of ``obs_compact`` / the int16 form of ``step_host``: float32 [..., 8] heads in the reference's order."""
        o = torch.empty(tuple(obs16.shape[:-1]) + (8,), dtype=torch.float32, device=obs16.device)
        f = obs16.to(torch.float32)
        o[..., 0], o[..., 1], o[..., 2], o[..., 3] = f[..., 0], 1.0, f[..., 1], f[..., 2]
        o[..., 4], o[..., 5], o[..., 6], o[..., 7] = float(self.config.width), float(self.config.height), f[..., 3], f[..., 4]
        return o

    def wait_host(self):
        """Block until the copies queued by ``step_host(..., wait=False)`` have completed."""
        _lib.check(self._lib.ofb_host_wait(self._h))

    def algorithmic_step_bytes(self):
        """SURVEY 8(d) byte model of one K1 launch: 88 B per ship + 64 B per live laser."""
        live = int(self.state(("n_lasers",))["n_lasers"].sum().item())
        return 88 * self.n_arenas * self.ships_number + 64 * live

    def frame(self, maps=None):
        """lib/battleground.py:163-166.  Without ``maps`` the raster (absolute_state) is produced on demand by
        ``raster()``; with ``maps`` (the "bits" buffer of ``raster``) the frame writes the new Observation's
        ship_map / laser_map there in the same launch (ofb_frame_bots / ofb_frame).  With the built-in device
        bots request_actions + generate_frame (+ the maps) is one fused launch: non-external ships draw their
        action inside the kernel, external ships (policy, host bots) use the rows already in ``self.actions``."""
        if isinstance(self.bot, ScriptedBots):
            self.bot._ensure_kinds(self)
            if maps is not None:
                _lib.check(self._lib.ofb_frame_bots(self._h, 0, _ptr(self.bot._kinds_dev), self.bot.seed, self.arena0,
                                                    self.total_steps, _ptr(self.actions), _ptr(self.obs_vec),
                                                    self._maps_arg(maps), self._stream()))
            else:
                _lib.check(self._lib.ofb_step_bots(self._h, 0, _ptr(self.bot._kinds_dev), self.bot.seed, self.arena0,
                                                   self.total_steps, _ptr(self.actions), _ptr(self.obs_vec),
                                                   self._stream()))
            self.launch_count += 1
            self.time += 1
            self.total_steps += 1
            return
        self.actions = self.request_actions()
        self.generate_frame(self.actions, maps=maps)

    def tick(self):
        """One controller tick with the MAX_TIME rule of lib/ofighters.py:684-688: after
        ``max_time`` frames the tick restarts the episode instead of stepping."""
        if self.time >= self.config.max_time:
            self.restart()
            return False
        self.frame()
        return True

    def watch_scores(self, arena):
        """``agent.scores`` for one arena (agents/agent.py:59-64: ``Agent.reset`` appends the finished episode's score): from now
        on every ``restart`` records the scores of that arena's ships first.  Returns the (live) list of int [S] rows."""
        if not 0 <= arena < self.n_arenas:
            raise Exception("Invalid arena {} : the batch holds {} arenas.".format(arena, self.n_arenas))
        if not hasattr(self, "_watched"):
            self._watched = {}
        return self._watched.setdefault(int(arena), [])

    def restart(self, mask=None, spawn_xy=None):
        """lib/battleground.py:108-117.  ``spawn_xy`` int [N,S,2] replaces the randint draws;
        ``mask`` bool [N] restricts the reset to some arenas (extension for vectorised envs)."""
        if getattr(self, "_watched", None):
            sc = self.state(("ship_score",))["ship_score"]
            for a, rows in self._watched.items():
                if mask is None or bool(torch.as_tensor(mask)[a]):
                    rows.append(sc[a].cpu().tolist())
        self.episode += 1
        if spawn_xy is None:
            self._draw_spawn(self.episode)
        else:
            self._spawn.copy_(torch.as_tensor(spawn_xy).to(torch.int32).reshape(self._spawn.shape))
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(self._lib.ofb_reset(self._h, _ptr(m), _ptr(self._spawn), _ptr(self.stats), self._stream()))
        _lib.check(self._lib.ofb_obs_vec(self._h, _ptr(self.obs_vec), self._stream()))
        self.launch_count += 2
        if mask is None:
            self.time = 0

    # ------------------------------------------------------------------ observation maps (K2)
    def raster(self, fmt="bits", out=None):
        """Observation.analyse_battleground for all arenas.  fmt: "bits" -> uint32 [N,2,W*H/32]
        (int32 tensor, bit y*W+x), "bf16" -> [N,H,W,2] NHWC as Keras' predict takes, "u8" likewise."""
        N, W, H = self.n_arenas, self.config.width, self.config.height
        if fmt == "bits":
            shape, dt, code = (N, 2, W * H // 32), torch.int32, _lib.OFB_MAP_BITS
        elif fmt == "bf16":
            shape, dt, code = (N, H, W, 2), torch.bfloat16, _lib.OFB_MAP_BF16
        elif fmt == "u8":
            shape, dt, code = (N, H, W, 2), torch.uint8, _lib.OFB_MAP_U8
        else:
            raise Exception("unknown map format " + str(fmt))
        if out is None:
            out = torch.empty(shape, dtype=dt, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != dt or not out.is_contiguous():
            raise Exception("Invalid map buffer : expected {} {}.".format(shape, dt))
        _lib.check(self._lib.ofb_raster(self._h, _ptr(out), code, self._stream()))
        self.launch_count += 1
        return out

    @property
    def absolute_state(self):
        """lib/battleground.py:105,117,166: the Observation of the whole battleground built after every frame -- here the
        two bit maps of every arena plus the per-ship heads (``.maps`` int32 [N,2,W*H/32], ``.obs_vec`` float32 [N,S,8])."""
        return SimpleNamespace(maps=self.raster("bits"), obs_vec=self.obs_vec, dim=self.dim, time=self.time)

    def running_stats(self):
        """The current episode's [sum score, kills, deaths, shots, ships, arenas] so far (int64 [6] on the device), without
        ending the episode -- ``restart`` adds the same six numbers of the finished episode into ``self.stats``."""
        out = torch.zeros(6, dtype=torch.int64, device=self.device)
        _lib.check(self._lib.ofb_stats(self._h, _ptr(out), self._stream()))
        self.launch_count += 1
        return out

    # ------------------------------------------------------------------ state access
    def state(self, fields=None):
        """Export the arena state as a dict of device tensors (names of include/ofb.h)."""
        N, S, L = self.n_arenas, self.ships_number, self.laser_cap
        spec = {}
        for n in ("time", "n_lasers", "kills", "deaths", "shots", "overflow", "episode", "near_ties"):
            spec[n] = ((N,), torch.int32)
        for n in ("ship_x", "ship_y", "ship_px", "ship_py", "ship_hull", "ship_reward", "ship_score", "ship_steps"):
            spec[n] = ((N, S), torch.int32)
        spec["ship_alive"] = ((N, S), torch.uint8)
        for n in ("laser_x", "laser_y", "laser_dx", "laser_dy"):
            spec[n] = ((N, L), torch.float64)
        for n in ("laser_owner", "laser_destroyed"):
            spec[n] = ((N, L), torch.uint8)
        view = _lib.OfbStateView()
        out = {}
        for n, (shape, dt) in spec.items():
            if fields is not None and n not in fields:
                continue
            out[n] = torch.empty(shape, dtype=dt, device=self.device)
            setattr(view, n, out[n].data_ptr())
        _lib.check(self._lib.ofb_state_export(self._h, C.byref(view), self._stream()))
        return out

    def load_state(self, tensors):
        """Import (a subset of) the state; inverse of ``state()``."""
        view = _lib.OfbStateView()
        keep = []
        for n, t in tensors.items():
            t = t.to(self.device).contiguous()
            keep.append(t)
            setattr(view, n, t.data_ptr())
        _lib.check(self._lib.ofb_state_import(self._h, C.byref(view), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def check_overflow(self):
        """Fail loudly if any arena ran out of laser slots (results would diverge from the reference)."""
        ov = int(self.state(("overflow",))["overflow"].sum().item())
        if ov:
            raise _lib.OfbError("laser-slot overflow in %d spawns: raise ArenaConfig.laser_cap" % ov)

    def arena(self, k):
        """Read-only ``Battleground``-shaped snapshot of arena k (ships[i].body.x, .pointing,
        .state, .agent.score/.reward/.steps; lasers[j].body.x/.y, .state, .owner; dim; time) so
        that UI-style consumers of the reference (lib/ofighters.py:578-641) keep working."""
        st = {n: v[k].cpu() for n, v in self.state().items()}
        ships = []
        for i in range(self.ships_number):
            ships.append(SimpleNamespace(
                id=i + 1, body=SimpleNamespace(x=int(st["ship_x"][i]), y=int(st["ship_y"][i]), radius=8),
                pointing=SimpleNamespace(x=int(st["ship_px"][i]), y=int(st["ship_py"][i])),
                state="flying" if int(st["ship_alive"][i]) else "destroyed", hull=int(st["ship_hull"][i]),
                time=int(st["time"]), can_shoot=1, laser_speed=10, max_speed=8,
                agent=SimpleNamespace(score=int(st["ship_score"][i]), reward=int(st["ship_reward"][i]),
                                      steps=int(st["ship_steps"][i]), behavior=self.behaviors[i],
                                      episode=int(st["episode"]))))
        lasers = []
        for j in range(int(st["n_lasers"])):
            lasers.append(SimpleNamespace(
                body=SimpleNamespace(x=float(st["laser_x"][j]), y=float(st["laser_y"][j]), radius=2),
                state="destroyed" if int(st["laser_destroyed"][j]) else "flying",
                owner=ships[int(st["laser_owner"][j])], speed=10))
        return SimpleNamespace(ships=ships, lasers=lasers, dim=self.dim, time=int(st["time"]),
                               ships_number=self.ships_number, ships_map=self.ships_map)
