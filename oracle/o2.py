"""ctypes binding of oracle O2 (oracle/mrf_oracle.c).  TEST INFRASTRUCTURE ONLY -- see mrf_oracle.h.

Also holds the shared scenario-record helpers (the 44-double per-robot record) used by tests and
bench.py's CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmrf_oracle.so")

MAX_ROBOTS, DOF, NLINKS, ROBOT_IN = 8, 7, 8, 44
Q, QD, G0, W0, G1, W1, G2, W2, ANG, CON, RB = 0, 7, 14, 17, 18, 21, 22, 23, 24, 33, 37


class Config(C.Structure):
    _fields_ = [
        ("n_robots", C.c_int), ("mode", C.c_int), ("static_or_dyn", C.c_int), ("has_collision_links", C.c_int),
        ("dt", C.c_double), ("eps", C.c_double), ("jdot_sign", C.c_double), ("jdot_ref_sign", C.c_double),
        ("exec_scale", C.c_double),
        ("mount", (C.c_double * 16) * MAX_ROBOTS),
        ("limits", (C.c_double * 2) * DOF),
        ("r_robots", (C.c_double * NLINKS) * MAX_ROBOTS),
        ("link_mask", C.c_int * MAX_ROBOTS),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mrf_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "mrf_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        d = C.POINTER(C.c_double)
        cp = C.POINTER(Config)
        _lib.mrfo_config_default.argtypes = [cp, C.c_int]
        _lib.mrfo_kinematics.argtypes = [cp, C.c_int, d, d, d, d, d, d]
        _lib.mrfo_endeffector.argtypes = [cp, C.c_int, d, d, C.c_int, d, d]
        _lib.mrfo_action.argtypes = [cp, C.c_int, d, C.c_int, d, d, d, d, d, d]
        _lib.mrfo_rollout_jointspace.argtypes = [cp, d, C.c_int, d, d, d, d]
        _lib.mrfo_rollout_jointspace_static.argtypes = [cp, d, C.c_int, C.c_int, d, d, d, d, d, d]
        _lib.mrfo_rollout_cartesian.argtypes = [cp, C.c_int, d, C.c_int, d, d, d, C.c_int, d, d, d]
        _lib.mrfo_rollout_jointspace_batch.argtypes = [cp, d, C.c_long, C.c_int, d, d, d, d, C.c_int]
        _lib.mrfo_action_batch.argtypes = [cp, C.c_int, d, C.c_long, C.c_int, d, d, d, d, d, C.c_int]
        _lib.mrfo_spheres.argtypes = [cp, C.c_int, d, d, C.c_int, d, d, d, d]
        _lib.mrfo_point_action.argtypes = [cp, d, d, d, C.c_double, C.c_double, C.c_int, d, d, C.c_int, d, d, d, d, d]
        _lib.mrfo_rollout_rfcv_batch.argtypes = [cp, d, C.c_long, C.c_int, C.c_int, C.c_double, C.c_int, d, d, d, d, d,
                                                 C.c_int]
        _lib.mrfo_max_threads.restype = C.c_int
        _lib.mrfo_hw_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_config(n_robots: int, **kw) -> Config:
    c = Config()
    lib().mrfo_config_default(C.byref(c), n_robots)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def mount_of(cfg: Config, robot: int) -> np.ndarray:
    return np.array(cfg.mount[robot][:]).reshape(4, 4)


def kinematics(cfg, robot, q, qd):
    q, qd = _c(q), _c(qd)
    x, v, c, J = np.zeros((8, 3)), np.zeros((8, 3)), np.zeros((8, 3)), np.zeros((8, 3, 7))
    lib().mrfo_kinematics(C.byref(cfg), robot, _p(q), _p(qd), _p(x), _p(v), _p(c), _p(J))
    return x, v, c, J


def endeffector(cfg, robot, q, qd, use_jqd=False):
    q, qd = _c(q), _c(qd)
    x, v = np.zeros(3), np.zeros(3)
    lib().mrfo_endeffector(C.byref(cfg), robot, _p(q), _p(qd), int(use_jqd), _p(x), _p(v))
    return x, v


def action(cfg, robot, rec, xo, vo, ao, ro, want_diag=False):
    rec = _c(rec)
    xo, vo, ao, ro = _c(xo).reshape(-1, 3), _c(vo).reshape(-1, 3), _c(ao).reshape(-1, 3), _c(ro).reshape(-1)
    S = ro.shape[0]
    out = np.zeros(7)
    diag = np.zeros(49 * 2 + 7 * 4) if want_diag else None
    rc = lib().mrfo_action(C.byref(cfg), robot, _p(rec), S, _p(xo), _p(vo), _p(ao), _p(ro), _p(out), _p(diag))
    if rc:
        raise FloatingPointError("oracle: Cholesky failed")
    if want_diag:
        d = dict(M_g=diag[0:49].reshape(7, 7), M_f=diag[49:98].reshape(7, 7), f_g=diag[98:105],
                 fe_g=diag[105:112], f_f=diag[112:119], qdd=diag[119:126])
        return out, d
    return out


def rollout_jointspace(cfg, rec, N, n_threads=0):
    """rec: (R,44) or (B,R,44).  Returns qN, qdN (.., R, N, 7), avg_vel (.., R), x_ee (.., R, 3)."""
    rec = _c(rec)
    R = cfg.n_robots
    single = rec.ndim == 2
    rec3 = rec.reshape(-1, R, ROBOT_IN)
    B = rec3.shape[0]
    qN, qdN = np.zeros((B, R, N, 7)), np.zeros((B, R, N, 7))
    avg, xee = np.zeros((B, R)), np.zeros((B, R, 3))
    lib().mrfo_rollout_jointspace_batch(C.byref(cfg), _p(rec3), B, N, _p(qN), _p(qdN), _p(avg), _p(xee), n_threads)
    # scenarios whose metric lost positive definiteness come back as NaN (see mrf_oracle.c)
    if single:
        return qN[0], qdN[0], avg[0], xee[0]
    return qN, qdN, avg, xee


def rollout_jointspace_static(cfg, rec, N, xs, rs):
    """One scenario: rec (R,44), static spheres xs (R,S,3), rs (R,S) per robot -> qN, qdN (R,N,7), avg_vel (R,), x_ee (R,3)."""
    rec, xs, rs = _c(rec), _c(xs), _c(rs)
    R = cfg.n_robots
    S = rs.shape[1]
    qN, qdN, avg, xee = np.zeros((R, N, 7)), np.zeros((R, N, 7)), np.zeros(R), np.zeros((R, 3))
    rc = lib().mrfo_rollout_jointspace_static(C.byref(cfg), _p(rec), N, S, _p(xs), _p(rs), _p(qN), _p(qdN), _p(avg), _p(xee))
    if rc == 2:
        raise ValueError("too many static spheres")
    return qN, qdN, avg, xee


def set_collision_links(cfg, links_per_robot):
    """collision_links_nrs: per robot a list of link numbers 1..8 -> cfg.link_mask."""
    for r, links in enumerate(links_per_robot):
        cfg.link_mask[r] = sum(1 << (int(l) - 1) for l in set(links))
    return cfg


def rollout_jointspace_avg(cfg, rec, N, n_threads=0):
    """As rollout_jointspace but without trajectory outputs (the get_velocity_rollouts path)."""
    rec = _c(rec)
    R = cfg.n_robots
    rec3 = rec.reshape(-1, R, ROBOT_IN)
    B = rec3.shape[0]
    avg, xee = np.zeros((B, R)), np.zeros((B, R, 3))
    lib().mrfo_rollout_jointspace_batch(C.byref(cfg), _p(rec3), B, N, None, None, _p(avg), _p(xee), n_threads)
    return avg, xee


def rollout_rfcv(cfg, rec, N, est_robot=1, est_h=0.2, use_jqd=False, trajectories=False, n_threads=0):
    """RF-CV control-step rollout of a batch (B,R,44): goal estimate of `est_robot` (example_pandas_Jointspace.py:346-348;
    -1 = none) + coupled rollout, all inside one OpenMP loop of the C library.
    -> dict(avg_vel (B,R), x_ee (B,R,3), goal_est (B,3)[, qN, qdN (B,R,N,7)])."""
    rec = _c(rec)
    R = cfg.n_robots
    rec3 = rec.reshape(-1, R, ROBOT_IN)
    B = rec3.shape[0]
    o = dict(avg_vel=np.zeros((B, R)), x_ee=np.zeros((B, R, 3)), goal_est=rec3[:, min(max(est_robot, 0), R - 1), G0:G0 + 3].copy())
    if trajectories:
        o["qN"], o["qdN"] = np.zeros((B, R, N, 7)), np.zeros((B, R, N, 7))
    lib().mrfo_rollout_rfcv_batch(C.byref(cfg), _p(rec3), B, N, est_robot, est_h, int(use_jqd), _p(o.get("qN")),
                                  _p(o.get("qdN")), _p(o["avg_vel"]), _p(o["x_ee"]), _p(o["goal_est"]), n_threads)
    return o


def rollout_cartesian(cfg, robot, rec, xo, vo, ro, N):
    rec = _c(rec)
    xo, vo, ro = _c(xo).reshape(-1, 3), _c(vo).reshape(-1, 3), _c(ro).reshape(-1)
    qN, qdN, avg = np.zeros((N, 7)), np.zeros((N, 7)), np.zeros(1)
    lib().mrfo_rollout_cartesian(C.byref(cfg), robot, _p(rec), ro.shape[0], _p(xo), _p(vo), _p(ro), N, _p(qN),
                                 _p(qdN), _p(avg))
    return qN, qdN, float(avg[0])


def sphere_offsets_ref(n: int) -> np.ndarray:
    """Oracle-side restatement of create_simulation_manipulators.py:188-245 -> (8, n, 3) link-frame offsets."""
    length = [0.333, 0.2, 0.3164, 0.2, 0.3840, 0.2, 0.088, 0.2]
    off = np.zeros((8, n, 3))
    for idx in range(8):
        z_start = length[idx] if idx % 2 == 0 else length[idx] / 2     # 'linear' / 'rotational' alternate (:193)
        for i in range(n):
            tr = np.array([0.0, 0.0, -z_start + i * length[idx] / n])
            if idx == 7:          # env link 16 == panda_joint8
                if i == 1:
                    tr = np.array([0.03, 0.03, -z_start + (i + 1) * length[idx] / n])
                elif i == 2:
                    tr[0:2] = [-0.03, -0.03]
            if idx == 4 and i in (2, 3):   # env link 11 == panda_joint5
                tr[0:2] = [0.0, 0.02 if i == 2 else 0.06]
            off[idx, i] = tr
    return off


def spheres(cfg, robot, q, qd, off):
    """-> x, v_origin, v_sphere, each (8 n, 3)."""
    q, qd, off = _c(q), _c(qd), _c(off)
    n = off.shape[1]
    x, vo, vs = np.zeros((8 * n, 3)), np.zeros((8 * n, 3)), np.zeros((8 * n, 3))
    lib().mrfo_spheres(C.byref(cfg), robot, _p(q), _p(qd), n, _p(off), _p(x), _p(vo), _p(vs))
    return x, vo, vs


def obstacle_lists(cfg, q, qd, off, vel_mode=0, static_or_dyn=1):
    """q, qd (R,7) -> per ego robot the (8 n (R-1), 10) obstacle records assembled as
    example_pandas_Jointspace.py:400-412 (vel_mode 0) / utils_apply_fk.py:3-33 (vel_mode 1) do."""
    R = cfg.n_robots
    n = off.shape[1]
    per = [spheres(cfg, j, q[j], qd[j], off) for j in range(R)]
    out = []
    for i in range(R):
        rows = []
        for j in range(R):
            if j == i:
                continue
            x, vo, vs = per[j]
            o = np.zeros((8 * n, 10))
            o[:, 0:3] = x
            o[:, 3:6] = (vs if vel_mode else vo) * (1.0 if static_or_dyn else 0.0)
            o[:, 9] = np.repeat(np.array(cfg.r_robots[j][:]), n)
            rows.append(o)
        out.append(np.concatenate(rows))
    return out


def point_action(cfg, q, qd, goal, w_goal, r_body, xs=(), rs=(), xd=(), vd=(), ad=(), rd=()):
    """Point-mass fabric action (mode 'acc'); static spheres xs (Ss,3), rs; dynamic spheres xd, vd, ad (Sd,2), rd."""
    q, qd, goal = _c(q), _c(qd), _c(goal)
    xs, rs = _c(xs).reshape(-1, 3), _c(rs).reshape(-1)
    xd, vd, ad, rd = _c(xd).reshape(-1, 2), _c(vd).reshape(-1, 2), _c(ad).reshape(-1, 2), _c(rd).reshape(-1)
    out = np.zeros(3)
    lib().mrfo_point_action(C.byref(cfg), _p(q), _p(qd), _p(goal), float(w_goal), float(r_body), len(rs), _p(xs), _p(rs),
                            len(rd), _p(xd), _p(vd), _p(ad), _p(rd), _p(out))
    return out


def max_threads() -> int:
    return int(lib().mrfo_max_threads())


def hw_threads() -> int:
    """Cores the process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
    return max(1, min(int(lib().mrfo_hw_threads()), len(os.sched_getaffinity(0))))


# --------------------------------------------------------------------------------------------- #
# record helpers
# --------------------------------------------------------------------------------------------- #
ROT_PANDA = np.array([[0.0, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]])  # parameters_manipulators.py:121


def make_record(q, qd, x_goal_0, weight_goal_0=2.0, x_goal_1=(0.107, 0.0, 0.0), weight_goal_1=10.0,
                x_goal_2=np.pi / 4, weight_goal_2=1.0, angle_goal_1=ROT_PANDA, constraint_0=(0.0, 0.0, 1.0, -0.65),
                radius_body=0.08) -> np.ndarray:
    r = np.zeros(ROBOT_IN)
    r[Q:Q + 7] = q
    r[QD:QD + 7] = qd
    r[G0:G0 + 3] = x_goal_0
    r[W0] = weight_goal_0
    r[G1:G1 + 3] = x_goal_1
    r[W1] = weight_goal_1
    r[G2] = float(np.asarray(x_goal_2).reshape(-1)[0])
    r[W2] = weight_goal_2
    r[ANG:ANG + 9] = np.asarray(angle_goal_1, dtype=np.float64).reshape(9)
    r[CON:CON + 4] = constraint_0
    r[RB:RB + 6] = radius_body
    return r


def record_to_params(rec) -> dict:
    """The fabrics parameter dict (names as in SURVEY A6) for oracle O1."""
    p = dict(x_goal_0=rec[G0:G0 + 3], weight_goal_0=rec[W0], x_goal_1=rec[G1:G1 + 3], weight_goal_1=rec[W1],
             x_goal_2=rec[G2:G2 + 1], weight_goal_2=rec[W2], angle_goal_1=rec[ANG:ANG + 9].reshape(3, 3),
             constraint_0=rec[CON:CON + 4])
    for i, l in enumerate(range(3, 9)):
        p[f"radius_body_panda_link{l}"] = rec[RB + i]
    return p
