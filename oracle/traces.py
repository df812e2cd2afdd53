"""Trace recording shared by the reference runner and the restatements -- test infrastructure.

A *trace* is a dict of numpy arrays describing E episodes x T frames of one
arena driven by an action tape ``actions[E,T,S,4]`` and spawn list
``spawn[E,S,2]`` (episode 0 = construction, later ones = ``restart()``):

  obs_vec[E,T,S,8]      what each bot is shown when asked for frame t's action
  ship_{x,y,px,py,alive,hull,reward,score}[E,T,S]   after frame t
  n_lasers[E,T], laser_{x,y,owner,destroyed}[E,T,LMAX]   list after frame t (append order)
  kills/deaths[E,T]     events so far in the episode ((1,t)/(10,t) of battleground.py:42)
  map_crc[E,T,2]        crc32 of np.packbits(ship_map), np.packbits(laser_map)
  maps[E,len(map_frames),2,W*H/8]  full packed maps for the frames in map_frames
  final_scores[E,S]     agent.scores entries appended by the resets
"""
import zlib

import numpy as np

LMAX = 512


def make_tapes(seed, E, T, S, kind="random", W=400, H=400):
    """Synthetic action tapes + spawns (inputs, by the north-star's definition).

    kind = "random": the reference random bot's distribution (agents/agent.py:123-133)
           "stress": always shoot, thrust~B(1/2), pointing U{0..W-1} (qlearnIA_V2.py:317-321)
           "lattice": everything on a 10-pixel lattice, little thrust -> many exact
                      angle ties in enemy_on_trajectory (hard part 2)
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    spawn = rng.integers(0, W + 1, size=(E, S, 2))
    actions = np.zeros((E, T, S, 4), dtype=np.int64)
    if kind == "random":
        k = rng.integers(0, 3, size=(E, T, S))
        actions[..., 0] = k == 0
        actions[..., 1] = k == 1
        new_pt = rng.integers(0, W + 1, size=(E, T, S, 2))
        for e in range(E):
            cur = spawn[e].copy()
            for t in range(T):
                rep = k[e, t] == 2
                cur[rep] = new_pt[e, t][rep]
                actions[e, t, :, 2:] = cur
        # exercise the reset quirks: x==0 keeps the old coordinate, 400 is off-map
        if E > 1:
            spawn[1, 0, 0] = 0
            spawn[1, 1, 1] = 0
            spawn[1, 2] = (400, 400)
    elif kind == "stress":
        actions[..., 0] = 1
        actions[..., 1] = rng.integers(0, 2, size=(E, T, S))
        actions[..., 2:] = rng.integers(0, W, size=(E, T, S, 2))
    elif kind == "lattice":
        spawn = rng.integers(4, 36, size=(E, S, 2)) * 10
        actions[..., 0] = rng.random((E, T, S)) < 0.7
        actions[..., 1] = rng.random((E, T, S)) < 0.03
        actions[..., 2:] = rng.integers(0, 41, size=(E, T, S, 2)) * 10
    else:
        raise ValueError(kind)
    return spawn.astype(np.int64), actions


def empty_trace(E, T, S, W, H, map_frames):
    z = lambda *sh, dt=np.int32: np.zeros(sh, dtype=dt)
    return {
        "obs_vec": z(E, T, S, 8, dt=np.float64),
        "ship_x": z(E, T, S), "ship_y": z(E, T, S), "ship_px": z(E, T, S), "ship_py": z(E, T, S),
        "ship_alive": z(E, T, S, dt=np.uint8), "ship_hull": z(E, T, S),
        "ship_reward": z(E, T, S), "ship_score": z(E, T, S),
        "n_lasers": z(E, T), "kills": z(E, T), "deaths": z(E, T),
        "laser_x": z(E, T, LMAX, dt=np.float64), "laser_y": z(E, T, LMAX, dt=np.float64),
        "laser_owner": z(E, T, LMAX, dt=np.uint8), "laser_destroyed": z(E, T, LMAX, dt=np.uint8),
        "map_crc": z(E, T, 2, dt=np.uint32),
        "maps": z(E, len(map_frames), 2, W * H // 8, dt=np.uint8),
        "map_frames": np.array(map_frames, dtype=np.int32),
        "final_scores": z(E, S),
    }


def pack_map(m):
    return np.packbits(np.asarray(m) != 0, axis=None, bitorder="little")


def record_maps(tr, e, t, ship_map, laser_map):
    ps, pl = pack_map(ship_map), pack_map(laser_map)
    tr["map_crc"][e, t] = (zlib.crc32(ps.tobytes()), zlib.crc32(pl.tobytes()))
    mf = list(tr["map_frames"])
    if t in mf:
        tr["maps"][e, mf.index(t), 0] = ps
        tr["maps"][e, mf.index(t), 1] = pl


def run_reference(spawn, actions, vector_ships=(), W=400, H=400, map_frames=()):
    """Trace of the real reference (container only)."""
    from oracle import ref_shim
    E, T, S, _ = actions.shape
    tr = empty_trace(E, T, S, W, H, map_frames)
    tapes = [actions[:, :, i].reshape(E * T, 4) for i in range(S)]
    arena = ref_shim.ReferenceArena(S, spawn[0], tapes, vector_ships=vector_ships, width=W, height=H)
    bg = arena.bg
    ev0 = 0
    for e in range(E):
        if e > 0:
            arena.restart(spawn[e])
            ev0 = len(bg.last_x_time_rewards)
        for t in range(T):
            arena.frame()
            for i, ship in enumerate(bg.ships):
                vec, _done = arena.bots[i].seen[e * T + t]
                tr["obs_vec"][e, t, i] = vec
                tr["ship_x"][e, t, i] = ship.body.x
                tr["ship_y"][e, t, i] = ship.body.y
                tr["ship_px"][e, t, i] = ship.pointing.x
                tr["ship_py"][e, t, i] = ship.pointing.y
                tr["ship_alive"][e, t, i] = ship.is_playable()
                tr["ship_hull"][e, t, i] = ship.hull
                tr["ship_reward"][e, t, i] = ship.agent.reward
                tr["ship_score"][e, t, i] = ship.agent.score
            n = len(bg.lasers)
            assert n <= LMAX
            tr["n_lasers"][e, t] = n
            for k, l in enumerate(bg.lasers):
                tr["laser_x"][e, t, k] = l.body.x
                tr["laser_y"][e, t, k] = l.body.y
                tr["laser_owner"][e, t, k] = bg.ships.index(l.owner)
                tr["laser_destroyed"][e, t, k] = l.state == "destroyed"
            ev = bg.last_x_time_rewards[ev0:]
            tr["kills"][e, t] = sum(1 for v, _ in ev if v == 1)
            tr["deaths"][e, t] = sum(1 for v, _ in ev if v == 10)
            record_maps(tr, e, t, bg.absolute_state.ship_map, bg.absolute_state.laser_map)
    # one more reset so that the last episode's score lands in agent.scores
    arena.restart(spawn[0])
    for i, ship in enumerate(bg.ships):
        tr["final_scores"][:, i] = ship.agent.scores[:E]
    return tr


def run_py(spawn, actions, W=400, H=400, map_frames=()):
    """Same trace from the pure-Python restatement."""
    from oracle.step_py import ArenaPy
    E, T, S, _ = actions.shape
    tr = empty_trace(E, T, S, W, H, map_frames)
    a = ArenaPy(spawn[0], W, H)
    for e in range(E):
        if e > 0:
            a.restart(spawn[e])
        for t in range(T):
            tr["obs_vec"][e, t] = a.obs_vec()
            a.step(actions[e, t])
            d = a.export(LMAX)
            for k in ("ship_x", "ship_y", "ship_px", "ship_py", "ship_alive", "ship_hull",
                      "ship_reward", "ship_score", "laser_x", "laser_y", "laser_owner", "laser_destroyed"):
                tr[k][e, t] = d[k]
            tr["n_lasers"][e, t] = d["n_lasers"]
            tr["kills"][e, t] = d["kills"]
            tr["deaths"][e, t] = d["deaths"]
            sm, lm = a.maps()
            record_maps(tr, e, t, sm, lm)
    a.restart(spawn[0])
    for i in range(S):
        tr["final_scores"][:, i] = a.scores[i][:E]
    return tr


def compare(a, b, keys=None):
    """Return the list of keys whose arrays differ bit-for-bit."""
    bad = []
    for k in (keys or a.keys()):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(k)
    return bad
