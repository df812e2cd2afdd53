"""Bridge between one arena of a batch and the reference's per-ship bot interface.

The reference's bot seam is ``bot.play(obs) -> Action | None`` (agents/agent.py:19-37) where ``obs``
is an ``Observation`` (lib/observation.py:40-133) and ``Action`` carries ``.shoot``, ``.thrust`` and
``.pointing`` (lib/action.py:24-41).  ``ArenaView`` shows arena k of a ``BatchedBattleground`` to
such a bot -- same field names, shapes and dtypes -- and writes the returned action into that
ship's row of the batch's action tensor, so an unmodified reference bot can drive ships of one
selected arena while every other arena runs device bots (debug / drop-in check).

``observation_fields`` is the pure host part (numpy in, numpy out) so that it can be checked
against the reference's own ``Observation`` class where the reference is importable.
"""
from types import SimpleNamespace

import numpy as np


def observation_fields(head, ship_bits, laser_bits, alive, width=400, height=400):
    """Build the reference's Observation fields of one ship from the batch's formats.

    head       float32[8]   (reward, can_shoot, pointing_x, pointing_y, W, H, x, y)  lib/observation.py:113-123
    ship_bits  uint32[W*H/32], laser_bits likewise: bit y*W+x, LSB first (ofb_raster OFB_MAP_BITS)
    alive      bool         ship is playable (done = not playable, lib/observation.py:110)
    """
    def unpack(bits):
        b = np.unpackbits(np.ascontiguousarray(bits, dtype=np.uint32).view(np.uint8), bitorder="little")
        return b[:width * height].reshape(height, width).astype(np.float64)       # indexed [row = y, col = x]

    ship_map, laser_map = unpack(ship_bits), unpack(laser_bits)
    head = np.asarray(head, dtype=np.float64)
    vector = np.concatenate([head, ship_map.ravel(), laser_map.ravel()]).reshape(-1, 1)   # (320008, 1) float64
    return SimpleNamespace(
        vector=vector, ship_map=ship_map, laser_map=laser_map,
        reward=head[0], can_shoot=int(head[1]),
        pointing=SimpleNamespace(x=int(head[2]), y=int(head[3])),
        dim=SimpleNamespace(x=int(head[4]), y=int(head[5])),
        pos=SimpleNamespace(x=int(head[6]), y=int(head[7])),
        done=not bool(alive))


def action_row(action, current_pointing):
    """Reference ``Action`` (or None) -> (shoot, thrust, pointing_x, pointing_y) int16 row.  ``None`` (what a dead
    ship's agent returns, lib/ship.py:260-262) keeps the current pointing and does nothing."""
    if action is None:
        return np.array([0, 0, current_pointing[0], current_pointing[1]], dtype=np.int16)
    p = getattr(action, "pointing", None)
    px, py = (int(p.x), int(p.y)) if p is not None else current_pointing
    return np.array([1 if action.shoot else 0, 1 if action.thrust else 0, px, py], dtype=np.int16)


class ArenaView:
    """Arena k of a BatchedBattleground as seen by reference-style bots."""

    def __init__(self, bg, k):
        if not 0 <= k < bg.n_arenas:
            raise Exception("arena index out of range")
        self.bg, self.k = bg, int(k)

    def observations(self):
        """One Observation-shaped object per ship (index order), from the batch's current state."""
        bg, k = self.bg, self.k
        heads = bg.obs_vec[k].cpu().numpy()
        maps = bg.raster("bits")[k].cpu().numpy().view(np.uint32)
        alive = bg.state(("ship_alive",))["ship_alive"][k].cpu().numpy()
        return [observation_fields(heads[i], maps[0], maps[1], alive[i], bg.config.width, bg.config.height)
                for i in range(bg.ships_number)]

    def play(self, bots):
        """``bots``: {ship index: object with play(obs)}.  Shows each its observation and stores the returned
        actions in the batch's action tensor (rows of ships without a bot are left to the device bots)."""
        import torch
        obs = self.observations()
        for i, bot in bots.items():
            o = obs[i]
            act = bot.play(o)                 # called for wreckage too: Ship.get_action (lib/ship.py:253-262) asks the agent
            if o.done:                        # and only then discards the answer -- a learning bot sees its own death there
                act = None
            row = action_row(act, (o.pointing.x, o.pointing.y))
            self.bg.actions[self.k, i].copy_(torch.from_numpy(row))
        return obs
