"""Drive any batched engine through a golden scenario and collect a trace (tests only)."""
import glob
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STATE_KEYS = ("ship_x", "ship_y", "ship_px", "ship_py", "ship_alive", "ship_hull", "ship_reward", "ship_score")
LASER_KEYS = ("laser_x", "laser_y", "laser_owner", "laser_destroyed")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def run_engine(make_engine, gold, n_copies=1, check_maps=True):
    """``make_engine(spawn[N,S,2]) -> engine``; replays ``gold``'s tape on n_copies arenas.

    Engine protocol: obs_vec() -> f32[N,S,8]; step(actions i16[N,S,4]); reset(spawn i32[N,S,2]);
    arrays() -> dict of numpy [N,...] incl. n_lasers/kills/deaths; raster_bits() -> u32[N,2,W*H/32].
    Returns a list of mismatch strings (empty = bit-exact for every copy, frame and field).
    """
    spawn, actions = gold["spawn"], gold["actions"]
    E, T, S, _ = actions.shape
    rep = lambda x: np.repeat(np.asarray(x)[None], n_copies, axis=0)
    eng = make_engine(rep(spawn[0]))
    bad = []
    for e in range(E):
        if e > 0:
            eng.reset(rep(spawn[e]))
        for t in range(T):
            ov = eng.obs_vec()
            if not np.array_equal(ov, rep(gold["obs_vec"][e, t]).astype(np.float32)):
                bad.append("obs_vec e%d t%d" % (e, t))
            eng.step(rep(actions[e, t]))
            arr = eng.arrays()
            n = int(gold["n_lasers"][e, t])
            for k in ("n_lasers", "kills", "deaths"):
                if not np.array_equal(arr[k], rep(gold[k][e, t])):
                    bad.append("%s e%d t%d" % (k, e, t))
            for k in STATE_KEYS:
                if not np.array_equal(arr[k].astype(np.int64), rep(gold[k][e, t]).astype(np.int64)):
                    bad.append("%s e%d t%d" % (k, e, t))
            for k in LASER_KEYS:
                if not np.array_equal(arr[k][:, :n], rep(gold[k][e, t, :n])):
                    bad.append("%s e%d t%d" % (k, e, t))
            if check_maps:
                bits = eng.raster_bits()
                for c in range(2):
                    crc = np.array([zlib.crc32(bits[i, c].tobytes()) for i in range(n_copies)], dtype=np.uint32)
                    if not np.all(crc == gold["map_crc"][e, t, c]):
                        bad.append("map%d e%d t%d" % (c, e, t))
                mf = list(gold["map_frames"])
                if t in mf:
                    full = gold["maps"][e, mf.index(t)].view(np.uint32).reshape(2, -1)
                    if not np.array_equal(bits, rep(full)):
                        bad.append("fullmap e%d t%d" % (e, t))
            if len(bad) > 20:
                return bad
    return bad


class COracleEngine:
    def __init__(self, spawn, lcap=512):
        from oracle.step_c import ArenasC
        self.a = ArenasC(spawn, lcap=lcap)

    def obs_vec(self):
        return self.a.obs_vec()

    def step(self, actions):
        self.a.step(actions)

    def reset(self, spawn):
        self.a.reset(spawn)

    def arrays(self):
        return self.a.arr

    def raster_bits(self):
        return self.a.raster_bits()
