"""Generate tests/golden/*.npz by running the REAL reference (build container only).

    python -m oracle.gen_golden

Each file = inputs (spawn, actions) + the trace the unmodified reference
produced for them under oracle/ref_shim.py (see oracle/traces.py for the keys).
The reference has no golden vectors of its own; these
files are the pin for oracle/step_py.py, oracle/step_c.c and the CUDA path.
libm matters for the trajectory reward (atan2/atan): generated with glibc 2.39.
"""
import os
import platform

import numpy as np

from oracle import traces

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SCENARIOS = [
    # name, kind, S, E, T, seed, vector_ships, map_frames
    ("default_s7_seed0", "random", 7, 2, 200, 0, (0,), (0, 1, 17, 60, 199)),
    ("default_s7_seed1", "random", 7, 2, 200, 1, (), (0, 33, 120)),
    ("default_s7_seed2", "random", 7, 3, 200, 2, (0, 3), (5, 150)),
    ("stress_s32_seed0", "stress", 32, 1, 80, 0, (), (0, 5, 12, 20, 40)),
    ("stress_s32_seed1", "stress", 32, 2, 60, 1, (), (10,)),
    ("lattice_s7_seed0", "lattice", 7, 2, 150, 0, (), (3,)),
    ("lattice_s12_seed1", "lattice", 12, 2, 150, 1, (), (3,)),
    ("lattice_s7_seed2", "lattice", 7, 2, 150, 2, (1,), (100,)),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, kind, S, E, T, seed, vec, mf in SCENARIOS:
        spawn, actions = traces.make_tapes(seed, E, T, S, kind)
        tr = traces.run_reference(spawn, actions, vector_ships=vec, map_frames=mf)
        lmax = int(tr["n_lasers"].max())
        for k in ("laser_x", "laser_y", "laser_owner", "laser_destroyed"):
            tr[k] = tr[k][..., :max(lmax, 1)]
        tr["obs_vec"] = tr["obs_vec"].astype(np.float32)
        tr["spawn"] = spawn.astype(np.int32)
        tr["actions"] = actions.astype(np.int16)
        tr["meta_kind"] = np.array(kind)
        tr["meta_libc"] = np.array(" ".join(platform.libc_ver()))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **tr)
        print(name, "lasers max", lmax, "kills", tr["kills"][:, -1], "score sum", int(tr["final_scores"].sum()),
              "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
