"""Pure-Python restatement of one Ofighters arena step -- test infrastructure.

Follows SURVEY.md Appendix A rule by rule; every block cites the reference
lines it restates (paths relative to /root/reference/ofighters).  It uses the
same CPython/libm expressions as the reference (``**``, ``math.sqrt``,
``math.atan2``, float ``%``) so that, on the same libm, results are
bit-identical to the reference run under ``oracle/ref_shim.py``.  That
equality is what tests/test_oracle_vs_reference.py and the committed traces in
tests/golden/ pin.

Only small cases should go through this file (it is as slow as the
reference); the batched C restatement in step_c.c covers the large ones.
"""
import math

import numpy as np

from oracle.disk import disk

R_SHIP = 8          # lib/ship.py:43
R_LASER = 2         # lib/laser.py:23
SHIP_SPEED = 8      # lib/ship.py:24
LASER_SPEED = 10    # lib/ship.py:86 * lib/laser.py:13
REWARDS = {"death": 0, "kill": 0, "aim": 2, "trajectory": 1}   # agents/qlearnIA_V2.py:39-44


class LaserPy:
    __slots__ = ("x", "y", "fx", "fy", "px", "py", "owner", "destroyed", "time")

    def __init__(self, x, y, fx, fy, px, py, owner):
        self.x, self.y = x, y
        self.fx, self.fy = fx, fy
        self.px, self.py = px, py
        self.owner = owner
        self.destroyed = False
        self.time = 0


class ArenaPy:
    """One arena with S ships.  Actions are rows ``(shoot, thrust, px, py)``."""

    def __init__(self, spawn_xy, width=400, height=400, rewards=None, max_time=200):
        self.W, self.H = int(width), int(height)
        self.S = len(spawn_xy)
        self.rw = dict(REWARDS if rewards is None else rewards)
        self.max_time = max_time
        self.time = 0
        # lib/ship.py:36-60 : position, pointing = own position, hull 1, flying
        self.x = [int(p[0]) for p in spawn_xy]
        self.y = [int(p[1]) for p in spawn_xy]
        self.px = list(self.x)
        self.py = list(self.y)
        self.alive = [True] * self.S
        self.hull = [1] * self.S
        # agents/agent.py:23-29
        self.reward = [0] * self.S
        self.score = [0] * self.S
        self.steps = [0] * self.S
        self.total_steps = [0] * self.S
        self.episode = 0
        self.scores = [[] for _ in range(self.S)]
        self.lasers = []
        # lib/battleground.py:42 event log -> counters
        self.kills = 0
        self.deaths = 0
        self.shots = 0

    # ---- A1: observation head + score fold ------------------------------
    def obs_vec(self):
        """``[reward, can_shoot, px, py, W, H, x, y]`` per ship (lib/observation.py:101-123)."""
        out = np.zeros((self.S, 8), dtype=np.float64)
        for i in range(self.S):
            out[i] = (self.reward[i], 1, self.px[i], self.py[i], self.W, self.H, self.x[i], self.y[i])
        return out

    def _fold_scores(self):
        # agents/agent.py:66-74, called for dead ships too (lib/ship.py:260-262)
        for i in range(self.S):
            self.steps[i] += 1
            self.total_steps[i] += 1
            self.score[i] += self.reward[i]
            self.reward[i] = 0

    # ---- A7: prune (lib/ofighters.py:623-625,704-707) --------------------
    def prune(self):
        self.lasers = [l for l in self.lasers if not l.destroyed]

    # ---- A2..A4 ---------------------------------------------------------
    def step(self, actions):
        """One ``Battleground.frame()`` minus the raster (lib/battleground.py:163-166)."""
        self.prune()
        self._fold_scores()
        self.time += 1                                           # battleground.py:155
        for l in self.lasers:                                    # battleground.py:157-158
            self._laser_move(l)
        for i in range(self.S):                                  # battleground.py:159-160
            self._ship_move(i, actions[i])

    def _outside(self, x, y):                                    # battleground.py:125-126
        return (x < 0) or (y < 0) or (x >= self.W) or (y >= self.H)

    def _laser_move(self, l):                                    # lib/laser.py:36-62
        l.time += 1
        dX = l.px - l.fx
        dY = l.py - l.fy
        dist = math.sqrt(dX ** 2 + dY ** 2)
        if dist != 0:
            l.x += dX * LASER_SPEED / dist
            l.y += dY * LASER_SPEED / dist
        explode = False
        for s in range(self.S):
            if self.alive[s] and math.sqrt((l.x - self.x[s]) ** 2 + (l.y - self.y[s]) ** 2) <= R_LASER + R_SHIP:
                self.reward[l.owner] += self.rw["kill"]
                self.kills += 1
                self.hull[s] -= 1                                # lib/ship.py:127-131
                if self.hull[s] <= 0:
                    self.reward[s] += self.rw["death"]           # lib/ship.py:225-230
                    self.deaths += 1
                    self.alive[s] = False
                explode = True
        if explode or self._outside(l.x, l.y):
            l.destroyed = True

    def _ship_move(self, i, action):                             # lib/ship.py:303-339
        if action is None or not self.alive[i]:
            return
        shoot, thrust, px, py = (int(v) for v in action)
        self.px[i], self.py[i] = px, py
        if thrust:                                               # lib/ship.py:213-222
            dX = self.px[i] - self.x[i]
            dY = self.py[i] - self.y[i]
            dist = math.sqrt(dX ** 2 + dY ** 2)
            if dist != 0:
                dx = dX * SHIP_SPEED / dist
                dy = dY * SHIP_SPEED / dist
                self.x[i] = min(self.W - 1, max(0, int(self.x[i] + dx)))
                self.y[i] = min(self.H - 1, max(0, int(self.y[i] + dy)))
        if shoot:                                                # lib/ship.py:134-156
            self._shoot(i)

    def _shoot(self, i):
        x, y, px, py = self.x[i], self.y[i], self.px[i], self.py[i]
        in_r = R_SHIP + R_LASER                                  # lib/form.py:159-188
        dX = px - x
        dY = py - y
        dist = math.sqrt(dX ** 2 + dY ** 2)
        if dist == 0:
            return
        sx = int(x + dX * in_r / dist)
        sy = int(y + dY * in_r / dist)
        fx, fy = sx, sy
        if math.sqrt((x - px) ** 2 + (y - py) ** 2) <= R_SHIP + R_LASER:   # lib/ship.py:147-148
            fx, fy = x, y
        self.lasers.append(LaserPy(sx, sy, fx, fy, px, py, i))
        self.shots += 1
        aimed = False                                            # lib/ship.py:158-169
        traj = False                                             # lib/ship.py:171-177
        for j in range(self.S):
            if j == i or not self.alive[j]:
                continue
            if math.sqrt((self.x[j] - px) ** 2 + (self.y[j] - py) ** 2) <= R_SHIP:
                aimed = True
            if self._on_trajectory(i, j):
                traj = True
        if aimed:
            self.reward[i] += self.rw["aim"]
        if traj:
            self.reward[i] += self.rw["trajectory"]

    def _on_trajectory(self, i, j):                              # lib/ship.py:179-210
        two_pi = 2 * math.pi
        shooting = math.atan2(self.py[i] - self.y[i], self.px[i] - self.x[i]) + math.pi
        if not shooting:
            return False
        target = math.atan2(self.y[j] - self.y[i], self.x[j] - self.x[i]) + math.pi
        if not target:
            return False
        d = math.sqrt((self.x[i] - self.x[j]) ** 2 + (self.y[i] - self.y[j]) ** 2)
        ang = two_pi if d == 0 else math.atan(R_SHIP / d)       # lib/form.py:298-307
        sup = (target + ang) % two_pi                            # lib/form.py:24-31
        inf = (target + -ang) % two_pi
        return inf <= shooting <= sup

    # ---- A8: episode reset ----------------------------------------------
    def restart(self, spawn_xy):
        """lib/battleground.py:108-117 + lib/ship.py:92-106 + agents/agent.py:59-64."""
        self.time = 0
        self.lasers = []
        for i in range(self.S):
            self.steps[i] = 0
            self.scores[i].append(self.score[i])
            self.score[i] = 0
            self.px[i], self.py[i] = self.x[i], self.y[i]        # pointing = OLD position
            nx, ny = int(spawn_xy[i][0]), int(spawn_xy[i][1])
            self.x[i] = nx or self.x[i]                          # 0 is falsy -> keeps old
            self.y[i] = ny or self.y[i]
            self.alive[i] = True                                 # hull NOT restored
        self.episode += 1
        self.kills = self.deaths = self.shots = 0

    # ---- A5: raster -------------------------------------------------------
    def maps(self):
        """``(ship_map, laser_map)`` indexed [row=y, col=x] (lib/observation.py:79-95)."""
        ship_map = np.zeros((self.W, self.H), dtype=np.uint8)
        laser_map = np.zeros((self.W, self.H), dtype=np.uint8)
        for s in range(self.S):
            if self.alive[s]:
                rr, cc = disk((self.y[s], self.x[s]), R_SHIP, shape=ship_map.shape)
                ship_map[rr, cc] = 1
        for l in self.lasers:
            rr, cc = disk((l.y, l.x), R_LASER, shape=laser_map.shape)
            laser_map[rr, cc] = 1
        return ship_map, laser_map

    # ---- export in the common array format --------------------------------
    def export(self, lcap=None):
        n = len(self.lasers)
        lcap = n if lcap is None else lcap
        d = {
            "time": np.int32(self.time), "n_lasers": np.int32(n),
            "kills": np.int32(self.kills), "deaths": np.int32(self.deaths), "shots": np.int32(self.shots),
            "ship_x": np.array(self.x, np.int32), "ship_y": np.array(self.y, np.int32),
            "ship_px": np.array(self.px, np.int32), "ship_py": np.array(self.py, np.int32),
            "ship_alive": np.array(self.alive, np.uint8), "ship_hull": np.array(self.hull, np.int32),
            "ship_reward": np.array(self.reward, np.int32), "ship_score": np.array(self.score, np.int32),
            "ship_steps": np.array(self.steps, np.int32),
            "laser_x": np.zeros(lcap), "laser_y": np.zeros(lcap),
            "laser_owner": np.zeros(lcap, np.uint8), "laser_destroyed": np.zeros(lcap, np.uint8),
        }
        for k, l in enumerate(self.lasers):
            d["laser_x"][k], d["laser_y"][k] = l.x, l.y
            d["laser_owner"][k], d["laser_destroyed"][k] = l.owner, l.destroyed
        return d
