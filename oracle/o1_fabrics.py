"""Oracle O1 -- structure-mirroring float64 restatement of the fabric action.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in
``fabrics==0.9.5`` / ``forwardkinematics==1.2.3`` / ``casadi==3.5.5``
(``/root/reference/poetry.lock:447,551,86``), none of which is vendored or
installable offline, and the reference's own tests pin no numbers
(``examples/test_examples.py:20``).  This file restates the *published
algorithm* of fabrics 0.9.5 the way that library is organised -- Lagrangian ->
Euler-Lagrange spec, Geometry, WeightedGeometry, DifferentialMap pull-back,
energisation, speed-control damper -- with every derivative taken by automatic
differentiation (torch.func, float64) exactly where fabrics takes it with
``ca.jacobian``/``ca.gradient``.  It is deliberately slow and generic; the
closed-form oracle O2 (``oracle/mrf_oracle.c``) and the CUDA kernels are checked
against it.  Geometry / Finsler / potential strings are the reference's own
(``examples/example_pandas_Jointspace.py:84-90``,
``examples/example_pointmasses_static.py:106-107``) or the fabrics 0.9.5 config
defaults, and are ``eval``'d against a tiny ``ca`` shim just as fabrics
``eval``s them against casadi.

Call sites this follows in the reference:
  planner construction  examples/example_pandas_Jointspace.py:64-134
  action call           examples/example_pandas_Jointspace.py:417-445
  joint-space rollout   multi_robot_fabrics/fabrics_planner/forward_planner_Jointspace.py:72-116,118-296
  Cartesian rollout     multi_robot_fabrics/fabrics_planner/forward_planner_Cartesian.py:347-489
  kinematics helpers    multi_robot_fabrics/utils/utils.py:16-54  (Jdot_sign = -1)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
from torch.func import grad, jacrev, jvp, vmap

F64 = torch.float64


# --------------------------------------------------------------------------- #
# casadi shim: the names the reference's strings use, mapped onto torch.
# sign()/heaviside() have zero derivative, as in CasADi.
# --------------------------------------------------------------------------- #
class _Ca:
    @staticmethod
    def sign(x):
        return torch.sign(x)   # zero derivative, as CasADi's sign

    @staticmethod
    def heaviside(x):
        one = torch.ones_like(x)   # piecewise constant: zero derivative, heaviside(0) = 0.5 as in CasADi
        return torch.where(x > 0, one, torch.where(x < 0, 0.0 * one, 0.5 * one))

    @staticmethod
    def exp(x):
        return torch.exp(x)

    @staticmethod
    def log(x):
        return torch.log(x)

    @staticmethod
    def tanh(x):
        return torch.tanh(x)

    @staticmethod
    def dot(a, b):
        return (a * b).sum()

    @staticmethod
    def norm_2(x):
        # CasADi SX simplifies sqrt(sq(x)) of a 1x1 to fabs(x) (derivative sign(x), 0 at 0); this is why
        # the reference does not produce NaN with q[6] == x_goal_2 == pi/4 at start
        # (parameters_manipulators.py:93, example_pandas_Jointspace.py:56).  For >1 element it is
        # sqrt(sumsqr(x)) and 0/0 = NaN at x == 0, as in the reference.
        if x.numel() == 1:
            return torch.abs(x).sum()
        return torch.sqrt((x * x).sum())

    @staticmethod
    def fmax(a, b):
        a = a if isinstance(a, torch.Tensor) else torch.tensor(float(a), dtype=F64)
        return torch.maximum(a, b)

    @staticmethod
    def SX(a):
        return torch.as_tensor(np.asarray(a), dtype=F64)


ca = _Ca()


def _eval(expr: str, **names):
    env = {"ca": ca, "np": np}
    env.update(names)
    return eval(expr, {"__builtins__": {}}, env)  # noqa: S307 (reference strings only)


# --------------------------------------------------------------------------- #
# fabrics 0.9.5 FabricPlannerConfig defaults (recalled)
# --------------------------------------------------------------------------- #
@dataclass
class FabricConfig:
    base_energy: str = "0.5 * 0.2 * ca.dot(xdot, xdot)"
    collision_geometry: str = "-0.5 / (x ** 5) * (-0.5 * (ca.sign(xdot) - 1)) * xdot ** 2"
    collision_finsler: str = "0.1/(x**2) * (-0.5 * (ca.sign(xdot) - 1)) * xdot**2"
    limit_geometry: str = "-0.1 / (x ** 1) * xdot ** 2"
    limit_finsler: str = "0.1/(x**1) * (-0.5 * (ca.sign(xdot) - 1)) * xdot**2"
    geometry_plane_constraint: str = "-0.5 / (x ** 5) * (-0.5 * (ca.sign(xdot) - 1)) * xdot ** 2"
    finsler_plane_constraint: str = "0.1/(x**2) * (-0.5 * (ca.sign(xdot) - 1)) * xdot**2"
    attractor_potential: str = "5.0 * (ca.norm_2(x) + 1 / 10 * ca.log(1 + ca.exp(-2 * 10 * ca.norm_2(x))))"
    attractor_metric: str = "((2.0 - 0.3) * ca.exp(-1 * (0.75 * ca.norm_2(x))**2) + 0.3) * ca.SX(np.identity(x.size()[0]))"
    damper_beta: str = "0.5 * (ca.tanh(-0.5 * (ca.norm_2(x) - 0.02)) + 1) * 6.5 + 0.01 + ca.fmax(0, a_ex - a_le)"
    damper_eta: str = "0.5 * (ca.tanh(-0.9 * (1 - 1/2) * ca.dot(xdot, xdot) - 0.5) + 1)"
    # restatement knobs (assumptions A7/A10-A12 of SURVEY Appendix A)
    eps: float = 1e-6
    jdot_sign: float = -1.0          # fabrics DifferentialMap default Jdot_sign
    exec_energy_scale: float = 1.0   # ExecutionLagrangian = scale * qdot.qdot


# The reference's Panda planner (examples/example_pandas_Jointspace.py:84-90)
def panda_config(**kw) -> FabricConfig:
    return FabricConfig(
        geometry_plane_constraint="10*(1/(1+1*ca.exp(-10*x))-1) * (xdot**2)",
        collision_geometry="-0.5 / (x ** 4) * (xdot ** 2)",
        collision_finsler="0.01/(x**4) * xdot**2",
        **kw,
    )


# The reference's point-mass planner (examples/example_pointmasses_static.py:106-107)
def pointmass_config(**kw) -> FabricConfig:
    return FabricConfig(
        collision_geometry="-2.0 / (x ** 1) * xdot ** 2",
        collision_finsler="1.0/(x**2) * (1 - ca.heaviside(xdot))* xdot**2",
        **kw,
    )


# --------------------------------------------------------------------------- #
# Kinematics: forwardkinematics GenericURDFFk restated for the Panda chain
# (URDF joints: examples/simulation_environments/urdfs/panda_with_finger.urdf:98-470)
# --------------------------------------------------------------------------- #
_HP = math.pi / 2
PANDA_JOINT_ORIGINS = [  # (xyz, rpy) of panda_joint1..7, all axes +z
    ((0.0, 0.0, 0.333), (0.0, 0.0, 0.0)),
    ((0.0, 0.0, 0.0), (-_HP, 0.0, 0.0)),
    ((0.0, -0.316, 0.0), (_HP, 0.0, 0.0)),
    ((0.0825, 0.0, 0.0), (_HP, 0.0, 0.0)),
    ((-0.0825, 0.384, 0.0), (-_HP, 0.0, 0.0)),
    ((0.0, 0.0, 0.0), (_HP, 0.0, 0.0)),
    ((0.088, 0.0, 0.0), (_HP, 0.0, 0.0)),
]
PANDA_LINK8_OFFSET = (0.0, 0.0, 0.107)  # fixed panda_joint8; panda_hand origin == link8 origin
PANDA_LIMITS = [  # examples/example_pandas_Jointspace.py:97-105
    [-2.8973, 2.8973], [-1.7628, 1.7628], [-2.8973, 2.8973], [-3.0718, -0.0698],
    [-2.8973, 2.8973], [-0.0175, 3.7525], [-2.8973, 2.8973],
]


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _hom(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def mount_transform(yaw: float, xyz) -> np.ndarray:
    """examples/example_pandas_Jointspace.py:108-118 / parameters_manipulators.py:138-150."""
    T = np.eye(4)
    T[0:2, 0:2] = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
    T[0:3, 3] = xyz
    return T


def _rotz(a):
    c, s = torch.cos(a), torch.sin(a)
    z, o = torch.zeros_like(a), torch.ones_like(a)
    return torch.stack([torch.stack([c, -s, z, z]), torch.stack([s, c, z, z]),
                        torch.stack([z, z, o, z]), torch.stack([z, z, z, o])])


def panda_link_frames(q: torch.Tensor, mount: np.ndarray):
    """4x4 world frames of panda_link1..panda_link8 (list of 8 tensors)."""
    T = torch.as_tensor(mount, dtype=F64)
    frames = []
    for i, (xyz, rpy) in enumerate(PANDA_JOINT_ORIGINS):
        T = T @ torch.as_tensor(_hom(_rpy(*rpy), xyz), dtype=F64) @ _rotz(q[i])
        frames.append(T)
    frames.append(frames[-1] @ torch.as_tensor(_hom(np.eye(3), PANDA_LINK8_OFFSET), dtype=F64))
    return frames


def panda_fk(q, mount, link: str):
    """``planner.get_forward_kinematics(link)`` (position only)."""
    frames = panda_link_frames(q, mount)
    if link == "panda_hand":
        return frames[7][:3, 3]
    assert link.startswith("panda_link")
    return frames[int(link[len("panda_link"):]) - 1][:3, 3]


def pointrobot_fk(q, link: str = "base_link"):
    """pointRobot1.urdf:91-113: prismatic x (origin z=0.05), prismatic y, revolute theta."""
    return torch.stack([q[0], q[1], 0.05 + 0.0 * q[2]])


# --------------------------------------------------------------------------- #
# fabrics building blocks
# --------------------------------------------------------------------------- #
def apply_euler(L, x, xdot, refs=None):
    """fabrics Lagrangian.applyEuler: returns (M, f_e).

    L(x, xdot[, x_ref, xdot_ref]) -> scalar; refs = (x_ref, xdot_ref, xddot_ref).
    """
    r = () if refs is None else (refs[0], refs[1])
    dL_dx = grad(L, 0)
    dL_dxd = grad(L, 1)
    M = jacrev(dL_dxd, 1)(x, xdot, *r)
    F = jacrev(dL_dx, 1)(x, xdot, *r)            # d2L_dxdxdot
    f = F.T @ xdot - dL_dx(x, xdot, *r)
    if refs is not None:
        f = f + jacrev(dL_dxd, 3)(x, xdot, *r) @ refs[2]   # d2L_dxdot dxdot_ref * xddot_ref
        f = f + jacrev(dL_dxd, 2)(x, xdot, *r) @ refs[1]   # d2L_dxdot dx_ref    * xdot_ref
    return M, f


def diff_map(phi, q, qdot, sigma):
    """fabrics DifferentialMap: J, Jdot*qdot with Jdot = sigma * d(J qdot)/dq."""
    J = jacrev(phi)(q)
    Jdot = sigma * jacrev(lambda qq: jacrev(phi)(qq) @ qdot)(q)
    return J, Jdot @ qdot


def pull(M, f, J, Jdq):
    """fabrics Spec.pull."""
    return J.T @ M @ J, J.T @ (f + M @ Jdq)


@dataclass
class Summed:
    M: torch.Tensor
    f: torch.Tensor
    fe: torch.Tensor   # Euler-Lagrange force of the summed (pulled) energies

    def __iadd__(self, o):
        self.M = self.M + o[0]
        self.f = self.f + o[1]
        if len(o) > 2:
            self.fe = self.fe + o[2]
        return self


class O1Planner:
    """Generic fabrics ParameterizedFabricPlanner restatement."""

    def __init__(self, dof: int, fk, config: FabricConfig, mode: str = "acc", time_step: float | None = None):
        self.dof = dof
        self.fk = fk                      # fk(q, link_name) -> (3,) tensor
        self.c = config
        self.mode = mode
        self.dt = time_step
        self.collision_links: list[str] = []
        self.n_static = 0
        self.n_dyn = 0
        self.dyn_dim = 3
        self.n_plane = 0
        self.limits = None
        self.goals: list[dict] = []

    # planner.set_components(...)  (examples/example_pandas_Jointspace.py:123-132)
    def set_components(self, collision_links=(), goal=(), number_obstacles=0, number_dynamic_obstacles=0,
                       dynamic_obstacle_dimension=3, number_plane_constraints=0, limits=None, skip_links=()):
        # links whose FK does not depend on q are skipped by fabrics (evidence
        # parameters_manipulators.py:41-43, forward_planner_Jointspace.py:163-166)
        self.collision_links = [l for l in collision_links if l not in skip_links]
        self.n_static = number_obstacles
        self.n_dyn = number_dynamic_obstacles
        self.dyn_dim = dynamic_obstacle_dimension
        self.n_plane = number_plane_constraints
        self.limits = limits
        self.goals = list(goal)

    # -- leaves -------------------------------------------------------------
    def _leaf_1d(self, geo: str, fin: str, x, xdot):
        """Leaf WeightedGeometry in its own 1-D coordinates: M = d2L/dxdot2, f = M h."""
        L = lambda xx, xd: _eval(fin, x=xx, xdot=xd).sum()
        M, _ = apply_euler(L, x, xdot)
        h = _eval(geo, x=x, xdot=xdot)
        return M, M @ h, L

    def _static_leaf(self, q, qdot, phi, geo, fin):
        s = self.c.jdot_sign
        x = phi(q)
        J, Jdq = diff_map(phi, q, qdot, s)
        xdot = J @ qdot
        M, f, L = self._leaf_1d(geo, fin, x, xdot)
        Mq, fq = pull(M, f, J, Jdq)
        # energy: Lagrangian.pull substitutes x -> phi(q), xdot -> J qdot, then applyEuler in q
        Lr = lambda qq, qd: L(phi(qq), jvp(phi, (qq,), (qd,))[1])
        _, fe = apply_euler(Lr, q, qdot)
        return Mq, fq, fe

    def _dynamic_sphere_leaf(self, q, qdot, fkl, x_ref, xd_ref, xdd_ref, rho, geo, fin):
        """DynamicObstacleLeaf: geometry map o dynamic map o fk map."""
        s = self.c.jdot_sign
        g = lambda xr: (torch.sqrt((xr * xr).sum()) / rho - 1.0).reshape(1)
        x_fk = fkl(q)
        J3, Jdq3 = diff_map(fkl, q, qdot, s)
        x_rel = x_fk - x_ref
        xd_rel = J3 @ qdot - xd_ref
        J1, Jdq1 = diff_map(g, x_rel, xd_rel, s)
        x = g(x_rel)
        xdot = J1 @ xd_rel
        M, f, L = self._leaf_1d(geo, fin, x, xdot)
        M1, f1 = pull(M, f, J1, Jdq1)          # geometry map
        f2 = f1 - M1 @ xdd_ref                 # Spec.dynamic_pull
        Mq, fq = pull(M1, f2, J3, Jdq3)        # forward kinematics

        def Lr(qq, qd, xr, xrd):
            xf, xfd = jvp(fkl, (qq,), (qd,))
            rel, reld = xf - xr, xfd - xrd
            xx, xxd = jvp(g, (rel,), (reld,))
            return L(xx, xxd)

        _, fe = apply_euler(Lr, q, qdot, (x_ref, xd_ref, xdd_ref))
        return Mq, fq, fe

    def _attractor(self, q, qdot, phi, weight):
        s = self.c.jdot_sign
        x = phi(q)
        J, Jdq = diff_map(phi, q, qdot, s)
        xdot = J @ qdot
        psi = lambda xx: (weight * _eval(self.c.attractor_potential, x=xx)).sum()
        h = grad(psi)(x)
        L = lambda xx, xd: (xd @ (_eval(self.c.attractor_metric, x=xx) @ xd))
        M, _ = apply_euler(L, x, xdot)
        return pull(M, M @ h, J, Jdq), x

    # -- planner.compute_action without the small-action clamp ---------------
    def action_raw(self, q, qdot, p: dict):
        """p: x_goal_j, weight_goal_j, angle_goal_j, x_obst_i, radius_obst_i, x_obst_dynamic_i,
        xdot_obst_dynamic_i, xddot_obst_dynamic_i, radius_obst_dynamic_i, radius_body_<link>,
        constraint_j   (fabrics parameter names)."""
        c = self.c
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=F64).reshape(-1)
        q, qdot = t(q), t(qdot)
        n = self.dof
        eye = torch.eye(n, dtype=F64)
        # base geometry: h = 0, energy = base_energy
        Lb = lambda qq, qd: _eval(c.base_energy, x=qq, xdot=qd)
        Mb, feb = apply_euler(Lb, q, qdot)
        G = Summed(Mb, torch.zeros(n, dtype=F64), feb)

        for link in self.collision_links:
            fkl = lambda qq, link=link: self.fk(qq, link)
            rb = float(np.asarray(p[f"radius_body_{link}"]).reshape(-1)[0])
            for i in range(self.n_static):
                xo = t(p[f"x_obst_{i}"])
                rho = float(np.asarray(p[f"radius_obst_{i}"]).reshape(-1)[0]) + rb
                phi = lambda qq, xo=xo, rho=rho, fkl=fkl: (torch.sqrt(((fkl(qq) - xo) ** 2).sum()) / rho - 1.0).reshape(1)
                G += self._static_leaf(q, qdot, phi, c.collision_geometry, c.collision_finsler)
            if self.n_dyn:
                d = self.dyn_dim
                fkd = lambda qq, fkl=fkl, d=d: fkl(qq)[0:d]
                S = range(self.n_dyn)
                xr = torch.stack([t(p[f"x_obst_dynamic_{i}"])[0:d] for i in S])
                xrd = torch.stack([t(p[f"xdot_obst_dynamic_{i}"])[0:d] for i in S])
                xrdd = torch.stack([t(p[f"xddot_obst_dynamic_{i}"])[0:d] for i in S])
                rho = torch.stack([t(p[f"radius_obst_dynamic_{i}"])[0] + rb for i in S])
                leaf = lambda a, b, cc, r: self._dynamic_sphere_leaf(q, qdot, fkd, a, b, cc, r, c.collision_geometry,
                                                                     c.collision_finsler)
                Mq, fq, fe = vmap(leaf)(xr, xrd, xrdd, rho)   # one leaf per obstacle (batched for speed only)
                G += (Mq.sum(0), fq.sum(0), fe.sum(0))
            for j in range(self.n_plane):
                cn = t(p[f"constraint_{j}"])
                phi = lambda qq, cn=cn, fkl=fkl, rb=rb: (((cn[0:3] * fkl(qq)).sum() + cn[3]) / torch.sqrt((cn[0:3] ** 2).sum()) - rb).reshape(1)
                G += self._static_leaf(q, qdot, phi, c.geometry_plane_constraint, c.finsler_plane_constraint)
        if self.limits is not None:
            for i, (lo, hi) in enumerate(self.limits):
                G += self._static_leaf(q, qdot, lambda qq, i=i, lo=lo: (qq[i] - lo).reshape(1), c.limit_geometry, c.limit_finsler)
                G += self._static_leaf(q, qdot, lambda qq, i=i, hi=hi: (hi - qq[i]).reshape(1), c.limit_geometry, c.limit_finsler)

        # forced geometry = geometry + attractors (add_forcing_geometry)
        Mf, ff = G.M.clone(), G.f.clone()
        x_psi = None
        for j, sg in enumerate(self.goals):
            w = float(np.asarray(p[f"weight_goal_{j}"]).reshape(-1)[0])
            xg = t(p[f"x_goal_{j}"])
            idx = sg["indices"]
            if sg["type"] == "staticJointSpaceSubGoal":
                phi = lambda qq, idx=idx, xg=xg: qq[idx] - xg
            else:
                child, parent = sg["child_link"], sg["parent_link"]
                R = None
                if isinstance(sg.get("angle"), (list, tuple)) and len(sg["angle"]) == 4:
                    R = torch.as_tensor(np.asarray(p[f"angle_goal_{j}"], dtype=np.float64), dtype=F64).reshape(3, 3)

                def phi(qq, child=child, parent=parent, R=R, idx=idx, xg=xg):
                    fc = self.fk(qq, child)
                    fp = torch.zeros(3, dtype=F64) if parent == "world" else self.fk(qq, parent)
                    if R is not None:
                        fc, fp = R @ fc, R @ fp
                    return fc[idx] - fp[idx] - xg
            (Ma, fa), xa = self._attractor(q, qdot, phi, w)
            Mf, ff = Mf + Ma, ff + fa
            if sg.get("is_primary_goal", False):
                x_psi = xa

        e = c.eps
        h_g = torch.linalg.solve(G.M + e * eye, G.f)
        h_f = torch.linalg.solve(Mf + e * eye, ff)
        xdd_f = -h_f
        if x_psi is None:           # no goal -> fabrics executes the unforced geometry
            qdd = -h_g
        else:
            a_geom = -(qdot @ (G.f - G.fe)) / (e + qdot @ (G.M @ qdot))
            s2 = 2.0 * c.exec_energy_scale     # M_ex = d2(scale*qd.qd)/dqd2
            den = e + s2 * (qdot @ qdot)
            a_ex0 = -(qdot @ (s2 * h_g)) / den
            a_exf = -(qdot @ (s2 * h_f)) / den
            eta = _eval(c.damper_eta, xdot=qdot)
            a_ex = eta * a_ex0 + (1.0 - eta) * a_exf
            beta = _eval(c.damper_beta, x=x_psi, a_ex=-a_ex, a_le=-a_geom)
            qdd = xdd_f - (a_ex + beta) * qdot
        act = qdd if self.mode == "acc" else qdot + self.dt * qdd
        diag = dict(M_g=G.M, f_g=G.f, fe_g=G.fe, M_f=Mf, f_f=ff, qdd=qdd)
        return act.numpy().copy(), {k: v.numpy().copy() for k, v in diag.items()}

    def compute_action(self, **p):
        """planner.compute_action incl. the |action| < 1e-6 -> 0 clamp."""
        a, _ = self.action_raw(p["q"], p["qdot"], p)
        if np.linalg.norm(a) < 1e-6:
            a = a * 0.0
        return a


# --------------------------------------------------------------------------- #
# The reference's concrete planners
# --------------------------------------------------------------------------- #
PANDA_GOAL = [  # examples/example_pandas_Jointspace.py:25-62
    dict(type="staticSubGoal", is_primary_goal=True, indices=[0, 1, 2], parent_link="world", child_link="panda_hand"),
    dict(type="staticSubGoal", is_primary_goal=False, indices=[0, 1, 2], parent_link="panda_link7",
         child_link="panda_hand", angle=[-0.366, 0.0, 0.0, 0.3305]),
    dict(type="staticJointSpaceSubGoal", is_primary_goal=False, indices=[6]),
]
PANDA_COLLISION_LINKS = [f"panda_link{i}" for i in range(1, 9)]


def make_panda_planner(mount: np.ndarray, n_dyn: int, n_static: int = 0, collision_links=PANDA_COLLISION_LINKS,
                       config: FabricConfig | None = None, mode="vel", time_step=0.01) -> O1Planner:
    """set_planner_panda (examples/example_pandas_Jointspace.py:64-134)."""
    pl = O1Planner(7, lambda q, link: panda_fk(q, mount, link), config or panda_config(), mode, time_step)
    pl.set_components(collision_links=collision_links, goal=PANDA_GOAL, number_obstacles=n_static,
                      number_dynamic_obstacles=n_dyn, dynamic_obstacle_dimension=3, number_plane_constraints=1,
                      limits=PANDA_LIMITS, skip_links=("panda_link1", "panda_link2"))
    return pl


POINT_GOAL = [dict(type="staticSubGoal", is_primary_goal=True, indices=[0, 1], parent_link="world", child_link="base_link")]


def make_point_planner(n_static: int, n_dyn: int = 0, config: FabricConfig | None = None) -> O1Planner:
    """set_planner_point (examples/example_pointmasses_static.py:102-129, _dynamic.py:102-131)."""
    pl = O1Planner(3, lambda q, link: pointrobot_fk(q, link), config or pointmass_config(), "acc", None)
    pl.set_components(collision_links=["base_link"], goal=POINT_GOAL, number_obstacles=n_static,
                      number_dynamic_obstacles=n_dyn, dynamic_obstacle_dimension=2 if n_dyn else 3)
    return pl


# --------------------------------------------------------------------------- #
# UtilsKinematics.necessary_kinematics (multi_robot_fabrics/utils/utils.py:16-54)
# --------------------------------------------------------------------------- #
def link_kinematics(q, qdot, mount, link: str, jdot_ref_sign: float = -1.0):
    """x = fk(q), v = J qdot, a = Jdot qdot with Jdot = jdot_ref_sign * d(J qdot)/dq (utils.py:28,37)."""
    t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=F64).reshape(-1)
    q, qdot = t(q), t(qdot)
    fkl = lambda qq: panda_fk(qq, mount, link)
    J, Jdq = diff_map(fkl, q, qdot, jdot_ref_sign)
    return fkl(q).numpy().copy(), (J @ qdot).numpy().copy(), Jdq.numpy().copy(), J.numpy().copy()


def sphere_kinematics(q, qdot, mount, link: str, t_link):
    """UtilsKinematics.define_symbolic_collision_link_poses (utils.py:99-114): position of the sphere
    fk(child=link, link_transformation=T)[0:3, 3] and its velocity jacobian(fk) @ qdot."""
    t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64), dtype=F64).reshape(-1)
    q, qdot, tl = t(q), t(qdot), t(t_link)
    idx = int(link[len("panda_link"):]) - 1

    def fks(qq):
        T = panda_link_frames(qq, mount)[idx]
        return T[:3, 3] + T[:3, :3] @ tl

    J = jacrev(fks)(q)
    return fks(q).numpy().copy(), (J @ qdot).numpy().copy()


# --------------------------------------------------------------------------- #
# Rollouts
# --------------------------------------------------------------------------- #
def jointspace_rollout(planners, mounts, q0, qd0, params, N, dt=0.01, r_robots=None, static_or_dyn=1):
    """ForwardFabricsPlanner.forward_multi_fabrics_symbolic (forward_planner_Jointspace.py:190-249),
    'vel' mode.  params[i] holds robot i's goal/constraint/radius_body parameters.
    Returns q_N, qdot_N (R x N x 7) and avg_vel (R,) (compute_velocity_average :102-116)."""
    R = len(planners)
    q = [np.asarray(x, dtype=np.float64).copy() for x in q0]
    qd = [np.asarray(x, dtype=np.float64).copy() for x in qd0]
    qN = np.zeros((R, N, 7))
    qdN = np.zeros((R, N, 7))
    for k in range(N):
        kin = []
        for i in range(R):                         # Phase A (:191-209)
            q[i] = q[i] + dt * qd[i]               # system_step, vel mode (:78-79)
            kin.append([link_kinematics(q[i], qd[i], mounts[i], l)[:3] for l in PANDA_COLLISION_LINKS])
        new_qd = []
        for i in range(R):                         # Phase B (:211-249)
            p = dict(params[i])
            o = 0
            for j in range(R):
                if j == i:
                    continue
                for l in range(8):
                    x, v, a = kin[j][l]
                    if static_or_dyn == 0:
                        v, a = np.zeros(3), np.zeros(3)
                    p[f"x_obst_dynamic_{o}"] = x
                    p[f"xdot_obst_dynamic_{o}"] = v
                    p[f"xddot_obst_dynamic_{o}"] = a
                    p[f"radius_obst_dynamic_{o}"] = (r_robots[j][l] if r_robots is not None else 0.08)
                    o += 1
            a, _ = planners[i].action_raw(q[i], qd[i], p)
            new_qd.append(a)
        for i in range(R):
            qd[i] = new_qd[i]
            qN[i, k] = q[i]
            qdN[i, k] = qd[i]
    avg = (qdN ** 2).sum(axis=(1, 2)) / (N * 7)
    return qN, qdN, avg


def cartesian_rollout(planner, q0, qd0, params, x_dyn, v_dyn, N, dt=0.01):
    """FabricsRollouts.symbolic_forward_fabrics (forward_planner_Cartesian.py:421-458), 'vel' mode."""
    q = np.asarray(q0, dtype=np.float64).copy()
    qd = np.asarray(qd0, dtype=np.float64).copy()
    x_dyn = [np.asarray(x, dtype=np.float64).copy() for x in x_dyn]
    qN = np.zeros((N, 7))
    qdN = np.zeros((N, 7))
    for k in range(N):
        p = dict(params)
        for o, (x, v) in enumerate(zip(x_dyn, v_dyn)):
            p[f"x_obst_dynamic_{o}"] = x
            p[f"xdot_obst_dynamic_{o}"] = np.asarray(v, dtype=np.float64)
            p[f"xddot_obst_dynamic_{o}"] = np.zeros(3)
        qd, _ = planner.action_raw(q, qd, p)
        q = q + dt * qd
        x_dyn = [x + dt * np.asarray(v, dtype=np.float64) for x, v in zip(x_dyn, v_dyn)]
        qN[k], qdN[k] = q, qd
    return qN, qdN, float((qdN ** 2).sum() / (N * 7))
