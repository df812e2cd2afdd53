"""Viewer bridge: one selected arena of a GPU-stepped batch, streamed back to the host in the shapes
the reference's Tk controller draws from, so that its drawing code keeps working on top of libofb:

    reference consumer                                          here
    ----------------------------------------------------------  ---------------------------------------------------
    Ofighters.actualise_ships / actualise_lasers                ViewerBridge.battleground  (ships[i].body.x/.y, .pointing,
      lib/ofighters.py:578-641 (reads battleground.ships,         .state, .agent.score ...; lasers[j].body, .state, .owner)
      .lasers, .dim, .time)
    ActionMapGraph.update  lib/action_map_graph.py:72-114       ViewerBridge.act_values (2,), .ptr_values (400, 400)
      (reads trainer.act_values, trainer.ptr_values)
    ScoreGraph / LossGraph / EpsilonGraph                        ViewerBridge.scores, .losses, .epsilons
      lib/score_graph.py, loss_graph.py, epsilon_graph.py
    Observation.ship_map / laser_map  (images/bot_inputs.png)    ViewerBridge.ship_map, .laser_map  (400, 400) float64

Only the selected arena crosses PCIe (a few KB of state + two 20 KB bit maps + one 640 KB pointer map per refresh); the batch
itself never leaves the GPU.  No compute happens here: the state comes from ofb_state_export, the maps from ofb_raster's bit
planes, the pointer map from ofb_policy_forward.
"""
from types import SimpleNamespace

import numpy as np
import torch


class ViewerBridge:
    def __init__(self, bg, arena=0, trainer=None, policy=None, ship=None, learner=None):
        if not 0 <= arena < bg.n_arenas:
            raise Exception("Invalid arena {} : the batch holds {} arenas.".format(arena, bg.n_arenas))
        self.bg = bg
        self.arena = int(arena)
        self.trainer = trainer
        self.policy = policy if policy is not None else (trainer.model if trainer is not None else None)
        ids = [i for i, b in enumerate(bg.behaviors) if b in ("QlearnIA", "external")]
        self.ship = int(ship) if ship is not None else (ids[0] if ids else 0)
        self.battleground = None
        self.ship_map = self.laser_map = None
        self.act_values = self.ptr_values = None
        self.learner = learner        # trainer.QLearner: the producer of .losses / .epsilons (QlearnIA.losses / .epsilons)
        # agent.scores of the selected arena's ships, one row per finished episode (agents/agent.py:59-64): the batch records them
        # at every restart for the arenas somebody watches
        self._scores = bg.watch_scores(self.arena)

    @property
    def scores(self):
        """``agent.scores`` of the selected ship: one entry per finished episode (what ScoreGraph plots)."""
        return [int(row[self.ship]) for row in self._scores]

    @property
    def epsilons(self):
        """``QlearnIA.epsilons`` (agents/qlearnIA_V2.py:364): epsilon at every reset, when a learner is attached."""
        return list(self.learner.epsilons) if self.learner is not None else []

    @property
    def losses(self):
        """``QlearnIA.losses`` (:382,408): one entry per replay."""
        return list(self.learner.losses) if self.learner is not None else []

    def refresh(self, maps_bits=None, with_pointer_map=True):
        """Pull the selected arena: state snapshot, the two observation maps and (when a policy is attached) the
        act / pointer values of its policy ship on the current observation."""
        bg, k = self.bg, self.arena
        self.battleground = bg.arena(k)
        if maps_bits is None:
            maps_bits = bg.raster("bits")
        b = maps_bits[k].cpu().numpy().view(np.uint32)
        dense = np.unpackbits(b.view(np.uint8).reshape(2, -1), axis=1, bitorder="little").reshape(2, bg.config.height, bg.config.width)
        self.ship_map, self.laser_map = dense[0].astype(np.float64), dense[1].astype(np.float64)   # [row = y, col = x]
        if self.policy is not None and with_pointer_map:
            r = self.policy.forward(maps_bits[k:k + 1].contiguous(), bg.obs_vec[k, self.ship].reshape(1, 8), 1, want_ptr=True)
            self.act_values = r["act"][0].cpu().numpy()
            self.ptr_values = r["ptr"][0].cpu().numpy()
            if self.trainer is not None:              # what ActionMapGraph reads (lib/action_map_graph.py:87-91)
                self.trainer.act_values, self.trainer.ptr_values = self.act_values, self.ptr_values
        return self

    def observation(self):
        """The selected ship's ``Observation`` fields (lib/observation.py:50-68) as plain host values."""
        v = self.bg.obs_vec[self.arena, self.ship].cpu().numpy()
        s = self.battleground.ships[self.ship] if self.battleground is not None else None
        return dict(reward=float(v[0]), can_shoot=float(v[1]), pointing=(int(v[2]), int(v[3])), dim=(int(v[4]), int(v[5])),
                    pos=(int(v[6]), int(v[7])), done=(s.state != "flying") if s is not None else None,
                    ship_map=self.ship_map, laser_map=self.laser_map)


class PlayerBridge:
    """Human ``Player`` passthrough (lib/player.py:6-63, Ship.read_keys lib/ship.py:233-242): the
    same event handlers as the reference's Player (without the Tk bindings -- bind them to any event source), and
    ``write_action`` = ``read_keys``: the pending keys become the int16 action row (shoot, thrust, pointing) of one
    "external" ship of one arena, the pointing staying where it was unless the cursor moved, then the keys are cleared."""

    def __init__(self, bg, arena=0, ship=0):
        if not 0 <= arena < bg.n_arenas or not 0 <= ship < bg.ships_number:
            raise Exception("Invalid arena / ship : ({}, {}).".format(arena, ship))
        if bg.behaviors[ship] not in ("external", "QlearnIA"):
            raise Exception("Ship {} is driven by the '{}' device bot: give the player an 'external' ship.".format(
                ship, bg.behaviors[ship]))
        self.bg, self.arena, self.ship = bg, int(arena), int(ship)
        self.shoot = False
        self.clear_keys()

    # ---- lib/player.py:33-63, event = anything with .x / .y for the cursor
    def press_shoot(self, event=None):
        self.actions_set.add("shoot")
        self.shoot = True

    def unpress_shoot(self, event=None):
        self.actions_set.discard("shoot")
        self.shoot = False

    def request_thrust(self, event=None):
        self.actions_set.add("thrust")
        self.thrust = True

    def request_turn(self, event):
        self.actions_set.add("pointing")
        self.turn = True
        self.cursor = SimpleNamespace(x=int(event.x), y=int(event.y))

    def clear_keys(self):
        self.actions_set = set()
        if self.shoot:
            self.actions_set.add("shoot")
        self.thrust = False
        self.turn = False
        self.cursor = None

    def write_action(self):
        """``Ship.read_keys`` (lib/ship.py:233-242) into ``bg.actions[arena, ship]``; call before ``bg.frame()``."""
        bg = self.bg
        shoot = "shoot" in self.actions_set
        thrust = "thrust" in self.actions_set
        if "pointing" in self.actions_set:
            px, py = self.cursor.x, self.cursor.y
        else:                                            # keep the ship's pointing (obs head: pointing_x, pointing_y)
            px, py = [int(v) for v in bg.obs_vec[self.arena, self.ship, 2:4].tolist()]
        row = torch.tensor([int(shoot), int(thrust), px, py], dtype=torch.int16, device=bg.device)
        bg.actions[self.arena, self.ship] = row
        self.clear_keys()
        return row
