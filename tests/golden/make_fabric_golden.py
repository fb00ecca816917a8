"""Generates tests/golden/fabric_golden.npz with oracle O1 (oracle/o1_fabrics.py: autodiff, fabrics-structured,
float64).  These vectors pin the closed-form oracle O2 and, through it, the CUDA kernels.  They are NOT outputs of
the reference itself (its arithmetic lives in casadi / fabrics wheels that cannot be installed offline -- "parity
unpinned"); they are an independent derivation of the same published algorithm.
Run in the build container (takes a few minutes):   python tests/golden/make_fabric_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import o1_fabrics as o1  # noqa: E402
from oracle import o2  # noqa: E402  (record helpers + default mounts only)


def params_with_obstacles(rec, obst):
    p = o2.record_to_params(rec)
    for i, o in enumerate(obst):
        p[f"x_obst_dynamic_{i}"] = o[0:3]
        p[f"xdot_obst_dynamic_{i}"] = o[3:6]
        p[f"xddot_obst_dynamic_{i}"] = o[6:9]
        p[f"radius_obst_dynamic_{i}"] = o[9]
    return p


def main():
    rng = np.random.default_rng(7)
    out = {}
    lim = np.array(o1.PANDA_LIMITS)
    cfg3 = o2.default_config(3)
    mounts = [o2.mount_of(cfg3, r) for r in range(3)]

    def rand_rec(w1=10.0, radius=0.08):
        q = rng.uniform(lim[:, 0] + 0.3, lim[:, 1] - 0.3)
        qd = rng.uniform(-0.5, 0.5, 7) * np.array([2.175] * 4 + [2.61] * 3)
        g = rng.uniform([0.2, -0.6, 0.8], [0.8, 0.6, 1.25])
        return o2.make_record(q, qd, g, rng.choice([2.0, 3.0]), weight_goal_1=w1, radius_body=radius)

    def rand_obst(S):
        o = np.zeros((S, 10))
        o[:, 0:3] = rng.uniform([-0.6, -1.2, 0.9], [1.6, 1.2, 2.0], size=(S, 3))
        o[:, 3:6] = rng.uniform(-0.3, 0.3, size=(S, 3))
        o[:, 6:9] = rng.uniform(-0.5, 0.5, size=(S, 3))
        o[:, 9] = 0.08
        return o

    # ---- single actions -------------------------------------------------------------------------------------
    cases = [dict(robot=0, S=3, mode="vel"), dict(robot=1, S=8, mode="vel"), dict(robot=2, S=1, mode="acc"),
             dict(robot=0, S=0, mode="vel", grasp=True), dict(robot=1, S=4, mode="vel", radii=True),
             dict(robot=0, S=2, mode="vel", zero_qd=True)]
    for c, cs in enumerate(cases):
        rec = rand_rec(w1=20.0)
        if cs.get("radii"):
            rec[o2.RB:o2.RB + 6] = [0.08, 0.07, 0.09, 0.06, 0.08, 0.1]
        if cs.get("zero_qd"):
            rec[o2.QD:o2.QD + 7] = 0.0
            rec[o2.QD + 2] = 0.3            # exercises sign(xdot) == 0 -> s = 0.5 in the limit / plane leaves
        obst = rand_obst(cs["S"])
        if cs.get("radii"):
            obst[:, 9] = rng.uniform(0.05, 0.12, size=cs["S"])
        links = [] if cs.get("grasp") else o1.PANDA_COLLISION_LINKS
        pl = o1.make_panda_planner(mounts[cs["robot"]], n_dyn=cs["S"], collision_links=links, mode=cs["mode"])
        a, d = pl.action_raw(rec[0:7], rec[7:14], params_with_obstacles(rec, obst))
        pre = f"act{c}_"
        out[pre + "robot"] = np.array(cs["robot"])
        out[pre + "mode"] = np.array(1 if cs["mode"] == "vel" else 0)
        out[pre + "grasp"] = np.array(1 if cs.get("grasp") else 0)
        out[pre + "rec"], out[pre + "obst"], out[pre + "action"] = rec, obst, a
        for k, v in d.items():
            out[pre + k] = v
        print("action case", c, a)
    out["n_act"] = np.array(len(cases))

    # ---- link kinematics (utils.py:16-54, Jdot_sign = -1) ------------------------------------------------------
    rec = rand_rec()
    kin = np.zeros((8, 3, 3))
    for l in range(8):
        x, v, a, _ = o1.link_kinematics(rec[0:7], rec[7:14], mounts[1], f"panda_link{l + 1}", jdot_ref_sign=-1.0)
        kin[l] = np.stack([x, v, a])
    out["kin_rec"], out["kin_robot"], out["kin_xva"] = rec, np.array(1), kin

    # ---- coupled joint-space rollout, 2 Pandas, N = 2 ----------------------------------------------------------
    R, N = 2, 2
    recs = np.stack([rand_rec(), rand_rec()])
    recs[0, 0:7] = [1.125, 0.19, 0.12, -1.66, 0.1, 1.88, 0.6]       # near the reference's pos0: arms face each other
    recs[1, 0:7] = [1.0, 0.3, 0.2, -1.5, -0.1, 1.9, 0.9]
    planners = [o1.make_panda_planner(mounts[r], n_dyn=8) for r in range(R)]
    qN, qdN, avg = o1.jointspace_rollout(planners, mounts[:R], [recs[r, 0:7] for r in range(R)],
                                         [recs[r, 7:14] for r in range(R)], [o2.record_to_params(recs[r]) for r in range(R)],
                                         N)
    out["ro_rec"], out["ro_N"], out["ro_qN"], out["ro_qdN"], out["ro_avg"] = recs, np.array(N), qN, qdN, avg
    print("rollout avg", avg)

    # ---- Cartesian rollout, N = 2, S = 2 -------------------------------------------------------------------------
    rec = rand_rec(w1=20.0)
    obst = rand_obst(2)
    pl = o1.make_panda_planner(mounts[0], n_dyn=2)
    p = o2.record_to_params(rec)
    for i in range(2):
        p[f"radius_obst_dynamic_{i}"] = obst[i, 9]
    qN, qdN, avg = o1.cartesian_rollout(pl, rec[0:7], rec[7:14], p, list(obst[:, 0:3]), list(obst[:, 3:6]), 2)
    out["cart_rec"], out["cart_obst"], out["cart_qN"], out["cart_qdN"], out["cart_avg"] = rec, obst, qN, qdN, np.array(avg)

    # ---- config C1: point masses (examples/example_pointmasses_static.py / _dynamic.py), oracle-only case ----------
    obstacles_pos = np.array([[1, 1.25, 0], [1, 3.75, 0], [1, -1.25, 0], [-1.1, 0, 0], [-1.1, 2.5, 0], [-1.1, -2.5, 0]], float)
    robots_pos = np.array([[-2.5, 0.01, 0.0], [-2.5, -2.49, 0.0], [2.5, 1.26, 0.0], [2.5, 3.74, 0.0]])
    goals = [np.array([1.5, 3.76]), np.array([1.5, 1.26]), np.array([-2.5, 0.01]), np.array([-2.5, -2.49])]
    vel = rng.uniform(-0.5, 0.5, size=(4, 3))
    pl_s = o1.make_point_planner(n_static=9)
    pl_d = o1.make_point_planner(n_static=6, n_dyn=3)
    pm_static, pm_dyn = [], []
    for i in range(4):
        others = [j for j in range(4) if j != i]
        p = dict(x_goal_0=goals[i], weight_goal_0=1.0, radius_body_base_link=0.2)
        for k in range(6):
            p[f"x_obst_{k}"], p[f"radius_obst_{k}"] = obstacles_pos[k], 1.0
        ps = dict(p)
        for k, j in enumerate(others):
            ps[f"x_obst_{6 + k}"], ps[f"radius_obst_{6 + k}"] = robots_pos[j], 0.2
        a, _ = pl_s.action_raw(robots_pos[i], vel[i], ps)
        pm_static.append(a)
        pd = dict(p)
        for k, j in enumerate(others):
            pd[f"x_obst_dynamic_{k}"], pd[f"xdot_obst_dynamic_{k}"] = robots_pos[j][0:2], vel[j][0:2]
            pd[f"xddot_obst_dynamic_{k}"], pd[f"radius_obst_dynamic_{k}"] = np.zeros(2), 0.2
        a, _ = pl_d.action_raw(robots_pos[i], vel[i], pd)
        pm_dyn.append(a)
    out["pm_pos"], out["pm_vel"], out["pm_goals"] = robots_pos, vel, np.array(goals)
    out["pm_obst"], out["pm_static"], out["pm_dyn"] = obstacles_pos, np.array(pm_static), np.array(pm_dyn)
    print("point-mass static", np.array(pm_static))

    # ---- coupled rollout, 3 Pandas, N = 2, STATIC_OR_DYN_FABRICS = 0 and unequal sphere radii (appended last so the
    #      random stream of the cases above is unchanged) ---------------------------------------------------------------
    R, N = 3, 2
    recs = np.stack([rand_rec(), rand_rec(), rand_rec()])
    recs[0, 0:7] = [1.1, -0.3, 0.0, -2.2, 0.0, 1.9, 0.8]          # near the reference's 3-robot pos0
    recs[1, 0:7] = [1.13, 0.2, 0.12, -1.65, 0.0, 1.86, 0.7]
    recs[2, 0:7] = [-0.47, -0.25, -0.4, -2.1, -0.1, 1.85, 0.37]
    recs[:, o2.RB:o2.RB + 6] = [0.08, 0.07, 0.09, 0.06, 0.08, 0.1]
    rr = [[0.08, 0.06, 0.08, 0.08, 0.07, 0.09, 0.08, 0.08], [0.05, 0.05, 0.08, 0.1, 0.08, 0.08, 0.06, 0.08],
          [0.08] * 8]
    planners = [o1.make_panda_planner(mounts[r], n_dyn=16) for r in range(R)]
    for sd in (1, 0):
        qN, qdN, avg = o1.jointspace_rollout(planners, mounts[:R], [recs[r, 0:7] for r in range(R)],
                                             [recs[r, 7:14] for r in range(R)],
                                             [o2.record_to_params(recs[r]) for r in range(R)], N, r_robots=rr,
                                             static_or_dyn=sd)
        out[f"ro3_sd{sd}_qN"], out[f"ro3_sd{sd}_qdN"], out[f"ro3_sd{sd}_avg"] = qN, qdN, avg
        print("3-robot rollout sd", sd, avg)
    out["ro3_rec"], out["ro3_rr"], out["ro3_N"] = recs, np.array(rr), np.array(N)

    np.savez_compressed(os.path.join(HERE, "fabric_golden.npz"), **out)
    print("wrote fabric_golden.npz")


if __name__ == "__main__":
    main()
