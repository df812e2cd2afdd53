"""ctypes front-end of oracle/step_c.c -- test infrastructure (checker / CPU baseline)."""
import ctypes as C

import numpy as np

from oracle import build as _build

BOT_KINDS = {"idle": 0, "random": 1, "turret": 2, "runner": 3, "thrust": 4, "shoot": 5, "stress": 6}

_I32 = ["time", "n_lasers", "kills", "deaths", "shots", "overflow", "episode"]
_SHIP_I32 = ["ship_x", "ship_y", "ship_px", "ship_py", "ship_hull", "ship_reward", "ship_score", "ship_steps"]
_LASER_F64 = ["laser_x", "laser_y"]
_LASER_I32 = ["laser_fx", "laser_fy", "laser_px", "laser_py"]
_LASER_U8 = ["laser_owner", "laser_destroyed"]


class _State(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_arenas", "n_ships", "lcap", "width", "height",
                                          "r_kill", "r_death", "r_aim", "r_traj")]
                + [(n, C.c_void_p) for n in _I32 + _SHIP_I32 + ["ship_alive"] + _LASER_F64
                   + _LASER_I32 + _LASER_U8 + ["stats"]])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.ofo_bot_actions.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p]
        _lib.ofo_random_spawn.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p]
        for f in ("ofo_step", "ofo_obs_vec", "ofo_reset", "ofo_raster_bits"):
            getattr(_lib, f).argtypes = None
    return _lib


class ArenasC:
    """N arenas x S ships stepped by the C restatement; arrays are numpy, owned here."""

    def __init__(self, spawn_xy, lcap=None, width=400, height=400,
                 rewards=(0, 0, 2, 1)):
        spawn_xy = np.asarray(spawn_xy, dtype=np.int32)
        N, S, _ = spawn_xy.shape
        self.N, self.S, self.W, self.H = N, S, width, height
        self.L = lcap or max(128, 16 * S)
        a = self.arr = {}
        for n in _I32:
            a[n] = np.zeros(N, np.int32)
        for n in _SHIP_I32:
            a[n] = np.zeros((N, S), np.int32)
        a["ship_alive"] = np.ones((N, S), np.uint8)
        for n in _LASER_F64:
            a[n] = np.zeros((N, self.L), np.float64)
        for n in _LASER_I32:
            a[n] = np.zeros((N, self.L), np.int32)
        for n in _LASER_U8:
            a[n] = np.zeros((N, self.L), np.uint8)
        a["stats"] = np.zeros(6, np.int64)
        a["ship_x"][:] = spawn_xy[..., 0]
        a["ship_y"][:] = spawn_xy[..., 1]
        a["ship_px"][:] = spawn_xy[..., 0]
        a["ship_py"][:] = spawn_xy[..., 1]
        a["ship_hull"][:] = 1
        st = self.st = _State()
        st.n_arenas, st.n_ships, st.lcap, st.width, st.height = N, S, self.L, width, height
        st.r_kill, st.r_death, st.r_aim, st.r_traj = rewards
        for n, v in a.items():
            setattr(st, n, v.ctypes.data)
        self._lib = lib()

    def __getattr__(self, k):
        arr = self.__dict__.get("arr", {})
        if k in arr:
            return arr[k]
        raise AttributeError(k)

    def step(self, actions, obs_out=None):
        actions = np.ascontiguousarray(actions, dtype=np.int16)
        assert actions.shape == (self.N, self.S, 4)
        self._lib.ofo_step(C.byref(self.st), C.c_void_p(actions.ctypes.data),
                           C.c_void_p(obs_out.ctypes.data if obs_out is not None else None))

    def obs_vec(self):
        out = np.zeros((self.N, self.S, 8), np.float32)
        self._lib.ofo_obs_vec(C.byref(self.st), C.c_void_p(out.ctypes.data))
        return out

    def reset(self, spawn_xy, mask=None):
        spawn = np.ascontiguousarray(spawn_xy, dtype=np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._lib.ofo_reset(C.byref(self.st), C.c_void_p(m.ctypes.data if m is not None else None),
                            C.c_void_p(spawn.ctypes.data))

    def raster_bits(self, out=None):
        if out is None:
            out = np.zeros((self.N, 2, self.W * self.H // 32), np.uint32)
        else:
            assert out.shape == (self.N, 2, self.W * self.H // 32) and out.dtype == np.uint32 and out.flags.c_contiguous
            out[...] = 0
        self._lib.ofo_raster_bits(C.byref(self.st), C.c_void_p(out.ctypes.data))
        return out

    def bot_actions(self, kind, seed, step, arena0=0):
        out = np.zeros((self.N, self.S, 4), np.int16)
        self._lib.ofo_bot_actions(C.byref(self.st), BOT_KINDS[kind], seed, arena0, step, out.ctypes.data)
        return out

    def random_spawn(self, seed, episode, arena0=0):
        out = np.zeros((self.N, self.S, 2), np.int32)
        self._lib.ofo_random_spawn(C.byref(self.st), seed, arena0, episode, out.ctypes.data)
        return out
