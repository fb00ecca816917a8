"""Drop-ins for the reference's kinematics helpers (caller-side glue of the hot path):

  UtilsKinematics.define_forward_kinematics        multi_robot_fabrics/utils/utils.py:16-85
  UtilsKinematics.define_symbolic_collision_link_poses   utils.py:87-119
  UtilsKinematics.define_symbolic_endeffector      utils.py:121-136
  compute_x_obsts_dyn_0, compute_endeffector       multi_robot_fabrics/utils/utils_apply_fk.py:3-44

The reference returns CasADi Functions; callers use them as ``f(q).full()``, ``jac(q) @ q_dot`` and
``jac_dot(q, q_dot) @ q_dot`` (examples/example_pandas_Jointspace.py:324-343,
fabrics_planner/forward_planner_Jointspace.py:90-99).  The objects here support exactly those uses; every number
comes from the CUDA kinematics / obstacle-staging kernels (B = 1 launches -- the batched entries in api.py are the
fast path).
"""
from __future__ import annotations

import numpy as np

from ._lib import DOF, MrfError, default_config
from .api import Fabrics


class _DM:
    """What callers do with a casadi.DM: ``.full()``, ``@``, numpy conversion."""

    def __init__(self, a):
        self._a = np.atleast_2d(np.asarray(a, dtype=np.float64))

    def full(self):
        return self._a

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __matmul__(self, other):
        return _DM(self._a @ np.asarray(other, dtype=np.float64).reshape(self._a.shape[1], -1))

    def __add__(self, other):
        return _DM(self._a + np.asarray(other, dtype=np.float64).reshape(self._a.shape))

    __radd__ = __add__


class _RobotKinematics:
    """One Fabrics handle per robot (its mount), evaluating all 8 links per call."""

    def __init__(self, mount, device=0):
        self.fab = Fabrics(config=default_config(1, mount=[np.asarray(mount, dtype=np.float64)], jdot_ref_sign=-1.0),
                           device=device)

    def xva(self, q, qd):
        x, v, a = self.fab.kinematics_host(np.asarray(q, dtype=np.float64).reshape(-1)[:DOF].reshape(1, 1, DOF),
                                           np.asarray(qd, dtype=np.float64).reshape(-1)[:DOF].reshape(1, 1, DOF))
        return x[0, 0], v[0, 0], a[0, 0]          # (8,3) each; a = Jdot_sign(-1) * d(J qd)/dq qd  (utils.py:28,37)

    def jacobians(self, q):
        """(8,3,7): column j = velocity of the link origins for a unit velocity of joint j (one batched launch)."""
        qb = np.broadcast_to(np.asarray(q, dtype=np.float64).reshape(-1)[:DOF].reshape(1, 1, DOF), (DOF, 1, DOF))
        _, v, _ = self.fab.kinematics_host(qb, np.eye(DOF).reshape(DOF, 1, DOF))
        return np.transpose(v[:, 0], (1, 2, 0))


class _JacDot:
    """``jac_dot_fun(q, qd) @ qd``: only the product is defined (that is the only way the reference uses it)."""

    def __init__(self, a):
        self._a = a

    def __matmul__(self, qd):
        return _DM(np.asarray(self._a).reshape(3, 1))


class UtilsKinematics:
    def __init__(self, device: int = 0):
        self.device = device
        self.nr_robots = 0
        self._kin = []

    def define_forward_kinematics(self, planners, collision_links_nrs, collision_links):
        """-> fk_dict with "fk_fun" / "jac_fun" / "jac_dot_fun" [robot][link] (utils.py:60-85)."""
        nr = len(collision_links_nrs)
        self.nr_robots = nr
        self._kin = [_RobotKinematics(planners[i].mount, self.device) for i in range(nr)]
        fk_dict = {k: [[] for _ in range(nr)] for k in ("fk_fun_center", "jac_fun_center", "jac_dot_fun_center", "fk_fun",
                                                        "jac_fun", "jac_dot_fun")}
        for i in range(nr):
            kin = self._kin[i]
            for link in collision_links[i]:
                l = (8 if link == "panda_hand" else int(link[len("panda_link"):])) - 1
                fk_dict["fk_fun"][i].append(lambda q, kin=kin, l=l: _DM(kin.xva(q, np.zeros(DOF))[0][l].reshape(3, 1)))
                fk_dict["jac_fun"][i].append(lambda q, kin=kin, l=l: _DM(kin.jacobians(q)[l]))
                fk_dict["jac_dot_fun"][i].append(lambda q, qd, kin=kin, l=l: _JacDot(kin.xva(q, qd)[2][l]))
        return fk_dict

    def define_symbolic_endeffector(self, planners):
        """-> [{"fk_fun_ee", "vel_fun_ee"}] per robot (utils.py:121-136): panda_hand position and J qdot."""
        if not self._kin:
            self._kin = [_RobotKinematics(p.mount, self.device) for p in planners]
            self.nr_robots = len(planners)
        out = []
        for kin in self._kin:
            out.append({"fk_fun_ee": lambda q, kin=kin: _DM(kin.xva(q, np.zeros(DOF))[0][7].reshape(3, 1)),
                        "vel_fun_ee": lambda q, qd, kin=kin: _DM(kin.xva(q, qd)[1][7].reshape(3, 1))})
        return out

    def define_symbolic_collision_link_poses(self, urdf_files, collision_links, sphere_transformations, n_obst_per_link=1,
                                             mount_transform=()):
        """-> [{"fk_fun": q8 -> (3, 8n), "vel_fun": (q8, qd8) -> (3, 8n)}] per robot (utils.py:87-119).  The 8-dof
        argument (7 joints + finger, utils_apply_fk.py:12-13) is accepted; the finger does not move a collision link."""
        import torch
        nr = len(sphere_transformations)
        self.nr_robots = nr
        out = []
        for i in range(nr):
            if any(l != f"panda_link{k + 1}" for k, l in enumerate(collision_links[i])):
                raise MrfError("the CUDA sphere staging implements the reference's link set panda_link1..8")
            off = np.array([[np.asarray(T)[0:3, 3] for T in link] for link in sphere_transformations[i]], dtype=np.float64)
            fab = Fabrics(config=default_config(1, mount=[np.asarray(mount_transform[i], dtype=np.float64)]),
                          device=self.device)

            def run(q, qd, fab=fab, off=off, want_v=False):
                dev = f"cuda:{self.device}"
                t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64).reshape(-1)[:DOF].reshape(DOF, 1, 1), device=dev)
                sx = torch.empty((8 * n_obst_per_link, 3, 1, 1), dtype=torch.float64, device=dev)
                sv = torch.empty_like(sx)
                fab.obstacles_dev(t(q), t(qd), n_per_link=n_obst_per_link, vel_mode=1, offsets=off, spheres_x=sx,
                                  spheres_v=sv, want_obst=False)
                return _DM((sv if want_v else sx)[:, :, 0, 0].T.cpu().numpy())

            out.append({"fk_fun": lambda q, run=run: run(q, np.zeros(DOF)),
                        "vel_fun": lambda q, qd, run=run: run(q, qd, want_v=True)})
        return out


def compute_x_obsts_dyn_0(q_robots, qdot_robots, x_collision_sphere_poses=None, nr_robots=2, fk_dict_spheres=(),
                          nr_dyn_obsts=(0, 0)):
    """Same call and return values as utils_apply_fk.py:3-33: per ego robot the other robots' sphere positions (taken
    from the environment dict, keys whose first element names the robot index) and sphere velocities (from the sphere
    functions, 8-dof argument = joints + one finger), and the per-robot position lists."""
    pad = lambda a: np.append(a, 0)
    own_x = {j: [x for key, x in x_collision_sphere_poses.items() if str(j) in key[0]] for j in range(nr_robots)}
    own_v = {j: fk_dict_spheres[j]["vel_fun"](pad(q_robots[j]), pad(qdot_robots[j])).full().transpose()
             for j in range(nr_robots)}
    x_dyn, v_dyn = [], []
    for ego in range(nr_robots):
        xs, vs = [], []
        for j in range(nr_robots):
            if j == ego:
                continue
            xs += own_x[j]
            vs.extend(np.vsplit(own_v[j], nr_dyn_obsts[ego]))
        x_dyn.append(xs)
        v_dyn.append(vs)
    return x_dyn, v_dyn, [own_x[j] for j in range(nr_robots)]


def compute_endeffector(q_robots, qdot_robots, fk_endeff, nr_robots=2):
    """utils_apply_fk.py:35-44."""
    x_ee, v_ee = [], []
    for i in range(nr_robots):
        x_ee.append(fk_endeff[i]["fk_fun_ee"](q_robots[i]).full().transpose()[0])
        v_ee.append(fk_endeff[i]["vel_fun_ee"](q_robots[i], qdot_robots[i]).full().transpose()[0])
    return x_ee, v_ee
