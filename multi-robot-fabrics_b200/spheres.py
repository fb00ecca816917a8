"""Collision-sphere placement along the Panda links for n_obst_per_link spheres per link (host-side constants).

Restates the placement rule of examples/simulation_environments/create_simulation_manipulators.py:188-245: per link a
"length" and a type (linear: spheres start one length below the link frame; rotational: half a length), sphere i at
z = -z_start + i * length / n in the link frame, plus the hand-tuned x/y shifts of the gripper link (urdf joint index 16,
panda_joint8) and of the bent link (index 11, panda_joint5).
"""
from __future__ import annotations

import numpy as np

LINK_LENGTH = [0.333, 0.2, 0.3164, 0.2, 0.3840, 0.2, 0.088, 0.2]          # :192
LINK_TYPE = ["linear", "rotational"] * 4                                   # :193


def sphere_offsets(n_obst_per_link: int) -> np.ndarray:
    """(8, n, 3) link-frame offsets of the spheres of panda_link1..8."""
    n = int(n_obst_per_link)
    off = np.zeros((8, n, 3))
    for l in range(8):
        length = LINK_LENGTH[l]
        z_start = length if LINK_TYPE[l] == "linear" else length / 2           # :213-220
        for i in range(n):
            z = -z_start + i * length / n                                       # :222
            t = [0.0, 0.0, z]
            if l == 7:                                                          # link == 16 (panda_joint8) :230-237
                if i == 1:
                    t = [0.03, 0.03, -z_start + (i + 1) * length / n]
                elif i == 2:
                    t = [-0.03, -0.03, z]
            if l == 4 and i in (2, 3):                                          # link == 11 (panda_joint5) :239-244
                t = [0.0, 0.02 if i == 2 else 0.06, z]
            off[l, i] = t
    return off
