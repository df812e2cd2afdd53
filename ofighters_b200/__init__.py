"""ofighters_b200 -- B200-native (sm_100a) implementation of Ofighters' data-parallel hot path:
batched arena step, observation raster and the bi-head policy forward, behind the reference's
own Battleground / bot / predict interfaces.  See DESIGN.md and include/ofb.h."""
from .config import ArenaConfig
from ._lib import OfbError

__all__ = ["ArenaConfig", "OfbError", "BatchedBattleground", "ScriptedBots"]


def __getattr__(name):
    if name in ("BatchedBattleground", "ScriptedBots"):
        from . import battleground
        return getattr(battleground, name)
    raise AttributeError(name)
