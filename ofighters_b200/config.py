"""Arena constants, mirrored from the reference's module-level constants."""
from dataclasses import dataclass


@dataclass(frozen=True)
class ArenaConfig:
    n_ships: int = 7            # lib/ofighters.py:53 SHIPS_MAP (6 + 1)
    laser_cap: int = 0          # 0 = auto: max(128, 16 * n_ships) slots per arena
    width: int = 400            # lib/observation.py:10
    height: int = 400           # lib/observation.py:11
    max_time: int = 200         # lib/ofighters.py:59 MAX_TIME
    reward_kill: int = 0        # agents/qlearnIA_V2.py:39-44 REWARDS
    reward_death: int = 0
    reward_aim: int = 2
    reward_trajectory: int = 1
    # fixed by the kernels (not configurable): ship radius 8 / hull 1 / speed 8 (lib/ship.py:43,45,24),
    # laser radius 2 / speed 10 (lib/laser.py:23, lib/ship.py:86)
