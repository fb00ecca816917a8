"""GPU tests of the reference-facing Python surface: the drop-in classes call the CUDA library and must reproduce
the oracle (fabric arithmetic) and the reference's own deadlock class (golden sequences)."""
import os
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200 import planner as P
from multi_robot_fabrics_b200.api import Fabrics
from oracle import o2
from oracle.deadlock_ref import DeadlockOracle

from helpers import oracle_episode, oracle_rollout, random_obstacles

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MOUNT = {"z_table": 0.65, "mount_positions": [np.array([0.0, 0.0, 0.65]), np.array([1.0, 0.0, 0.65]),
                                              np.array([0.7, 0.6, 0.65])]}
LINKS = [1, 2, 3, 4, 5, 6, 7, 8]


def make_params(R, N):
    """The attributes ForwardFabricsPlanner reads from manipulator_parameters (parameters_manipulators.py)."""
    p = types.SimpleNamespace()
    p.N_HORIZON, p.dt, p.dof, p.nr_obsts = N, 0.01, [7] * R, [0] * R
    p.fabrics_mode, p.STATIC_OR_DYN_FABRICS = "vel", 1
    p.collision_links_nrs = [LINKS] * R
    p.r_robots = [[0.08] * 8 for _ in range(R)]
    p.rotation_matrix_pandas = [P.ROT_PANDA] * R
    p.nr_obsts_dyn = [8 * (R - 1)] * R
    return p


def test_compute_action_dropin_matches_oracle(built):
    """planner.compute_action(**kwargs) with the kwargs of examples/example_pandas_Jointspace.py:421-439."""
    rng = np.random.default_rng(0)
    S = 8
    rec = m.scenarios.generate(6, 2, seed=51, weight_goal_1=20.0)
    obst = random_obstacles(rng, 6, 2, S, rec)
    ocfg = o2.default_config(2)
    for robot in (0, 1):
        pl, goal = P.set_planner_panda(7, 0, S, LINKS, {}, MOUNT, robot)
        assert len(goal._config) == 3
        for b in range(6):
            r, o = rec[b, robot], obst[b, robot]
            xs = [o[i, 0:3] for i in range(S)]
            kw = dict(q=r[0:7], qdot=r[7:14], x_goal_0=r[14:17], weight_goal_0=r[17], angle_goal_1=P.ROT_PANDA,
                      x_goal_1=np.array([0.107, 0.0, 0.0]), weight_goal_1=20.0, x_goal_2=np.array([np.pi / 4]),
                      weight_goal_2=1.0, x_obsts=xs, radius_obsts=[0.08] * S, constraint_0=np.array([0, 0, 1, -0.65]),
                      radius_body_panda_links={str(l): np.array(0.08) for l in range(3, 9)},
                      radius_body_panda_hand=np.array([0.08]), x_obsts_dynamic=xs,
                      xdot_obsts_dynamic=[o[i, 3:6] for i in range(S)], xddot_obsts_dynamic=[o[i, 6:9] for i in range(S)],
                      radius_obsts_dynamic=[0.08] * S)
            act = pl.compute_action(**kw)
            ref = o2.action(ocfg, robot, r, o[:, 0:3], o[:, 3:6], o[:, 6:9], o[:, 9])
            assert act.shape == (7,) and np.abs(act - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    # grasp planner (collision_links_nr=[]) and the small-action clamp
    pg, _ = P.set_planner_panda(7, 1, 1, [], {}, MOUNT, 1)
    r = rec[0, 1]
    act = pg.compute_action(q=r[0:7], qdot=r[7:14], x_goal_0=r[14:17], weight_goal_0=r[17], angle_goal_1=P.ROT_PANDA,
                            x_goal_1=[0.107, 0, 0], weight_goal_1=20.0, x_goal_2=[np.pi / 4], weight_goal_2=1.0,
                            x_obsts=[np.zeros(3)], radius_obsts=[0.1], constraint_0=[0, 0, 1, -0.65])
    ref = o2.action(o2.default_config(2, has_collision_links=0), 1, np.concatenate([r[:21], [20.0], r[22:]]),
                    np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    assert np.abs(act - ref).max() < 1e-9


def test_forward_fabrics_planner_dropin(built):
    R, N = 2, 20
    rec = m.scenarios.generate(3, R, seed=52)
    planners = [P.set_planner_panda(7, 0, 8, LINKS, {}, MOUNT, i)[0] for i in range(R)]
    goals = [P.PandaGoal() for _ in range(R)]
    fp = P.ForwardFabricsPlanner(make_params(R, N), planners, 100, None, goals)
    assert fp.forward_multi_fabrics_symbolic() == {}
    for b in range(3):
        ia = {"q_robots": [rec[b, i, 0:7] for i in range(R)], "q_dot_robots": [rec[b, i, 7:14] for i in range(R)],
              "x_obsts": [[] * R], "x_goals0": [rec[b, i, 14:17] for i in range(R)],
              "x_goals1": [goals[i]._config.subgoal1.desired_position for i in range(R)],
              "x_goals2": [goals[i]._config.subgoal2.desired_position for i in range(R)],
              "weight_goals0": [rec[b, i, 17] for i in range(R)],
              "weight_goals1": [goals[i]._config.subgoal1.weight for i in range(R)],
              "weight_goals2": [goals[i]._config.subgoal2.weight for i in range(R)],
              "constraints": [np.array([0, 0, 1, -0.65])] * R}
        vel_avg = fp.get_velocity_rollouts(ia)
        qN, qdN, avg, *_ = oracle_rollout(rec[b:b + 1], R, N)
        assert len(vel_avg) == R and vel_avg[0].shape == (1,)
        assert np.abs(np.array([v[0] for v in vel_avg]) - avg[0]).max() < 1e-9
        q_n, qd_n, qdd_n = fp.rollouts_numerical(ia)
        for i in range(R):
            assert q_n[f"robot_{i}"][0].shape == (7, N)
            assert np.abs(q_n[f"robot_{i}"][0] - qN[0, i].T).max() < 1e-9
            assert np.abs(qd_n[f"robot_{i}"][0] - qdN[0, i].T).max() < 1e-9
            assert not qdd_n[f"robot_{i}"][0].any()


def test_forward_fabrics_planner_acc_mode(built):
    """fabrics_mode 'acc' in the coupled joint-space rollouts: the reference's loop resets the acceleration it integrates to
    zero each step and stores the planner output in q_dot (forward_planner_Jointspace.py:195,202,233) -- the oracle with
    mode = 0 runs that very recurrence."""
    R, N, nb = 2, 3, 16
    rec = m.scenarios.generate(nb, R, seed=58)
    rec[:, :, 7:14] *= 0.1            # the recurrence of this mode diverges within a few steps: short horizon, slow start
    planners = [P.set_planner_panda(7, 0, 8, LINKS, {}, MOUNT, i)[0] for i in range(R)]
    goals = [P.PandaGoal() for _ in range(R)]
    prm = make_params(R, N)
    prm.fabrics_mode = "acc"
    fp = P.ForwardFabricsPlanner(prm, planners, 100, None, goals)
    ocfg = o2.default_config(R, mode=0)
    compared = 0
    for b in range(nb):
        ia = {"q_robots": [rec[b, i, 0:7] for i in range(R)], "q_dot_robots": [rec[b, i, 7:14] for i in range(R)],
              "x_obsts": [[] * R], "x_goals0": [rec[b, i, 14:17] for i in range(R)],
              "x_goals1": [goals[i]._config.subgoal1.desired_position for i in range(R)],
              "x_goals2": [goals[i]._config.subgoal2.desired_position for i in range(R)],
              "weight_goals0": [rec[b, i, 17] for i in range(R)],
              "weight_goals1": [goals[i]._config.subgoal1.weight for i in range(R)],
              "weight_goals2": [goals[i]._config.subgoal2.weight for i in range(R)],
              "constraints": [np.array([0, 0, 1, -0.65])] * R}
        qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec[b], N)
        if not np.isfinite(qdN).all():
            continue
        q_n, qd_n, _ = fp.rollouts_numerical(ia)
        for i in range(R):
            scale = max(1.0, np.abs(qdN[i]).max())
            assert np.abs(q_n[f"robot_{i}"][0] - qN[i].T).max() < 1e-9 * scale
            assert np.abs(qd_n[f"robot_{i}"][0] - qdN[i].T).max() < 1e-9 * scale
        vel_avg = fp.get_velocity_rollouts(ia)
        assert np.abs(np.array([v[0] for v in vel_avg]) - avg).max() < 1e-9 * max(1.0, np.abs(avg).max())
        compared += 1
    assert compared >= 3


def test_fabrics_rollouts_cartesian_dropin(built):
    N, S = 10, 8
    rng = np.random.default_rng(3)
    rec = m.scenarios.generate(2, 2, seed=53, weight_goal_1=20.0)
    obst = random_obstacles(rng, 2, 2, S, rec)
    ocfg = o2.default_config(2)
    for robot in (0, 1):
        pl, goal = P.set_planner_panda(7, 0, S, LINKS, {}, MOUNT, robot)
        fr = P.FabricsRollouts(N=N, dt=0.01, nx=7, nu=7, dof=7, nr_obsts=0, bool_ring=False, nr_obsts_dyn=S,
                               v_obsts_dyn=[np.zeros(3)] * S, fabrics_mode="vel", collision_links_nrs=LINKS,
                               nr_constraints=1, constraints=np.array([0, 0, 1, -0.65]), nr_goals=3)
        fr.preset_radii_obsts_dyn([0.08] * S)
        fr.symbolic_forward_fabrics(pl, goal)
        r, o = rec[0, robot], obst[0, robot]
        args = fr.define_arguments_numerical(
            q_robot=r[0:7], q_dot_robot=r[7:14], weight_goals={"subgoal0": r[17], "subgoal1": 20.0, "subgoal2": 1.0},
            x_goals={"subgoal0": r[14:17], "subgoal1": np.array([0.107, 0, 0]), "subgoal2": np.array([np.pi / 4])},
            x_obsts=[], x_obsts_dyn=[o[i, 0:3] for i in range(S)], v_obsts_dyn=[o[i, 3:6] for i in range(S)],
            constraints=np.array([0, 0, 1, -0.65]))
        q_n, qd_n, qdd_n = fr.rollouts_numerical(args)
        rq, rqd, ravg = o2.rollout_cartesian(ocfg, robot, r, o[:, 0:3], o[:, 3:6], o[:, 9], N)
        assert q_n.shape == (7, N) and np.abs(q_n - rq.T).max() < 1e-9 and np.abs(qd_n - rqd.T).max() < 1e-9
        assert abs(fr.get_velocity_rollouts(args).full()[0][0] - ravg) < 1e-9


def test_cartesian_python_rollout_loop_and_obstacle_helpers(built):
    """FabricsRollouts.forward_fabrics ("the Python rollout loop", forward_planner_Cartesian.py:218-273), get_action
    (:132-191), get_x_obsts_dyn_N (:195-216), x_obsts_dyn_numerical (:491-505) and system_step (:77-92): N actions through
    the CUDA action kernel reproduce the oracle's decoupled rollout, and the constant-velocity obstacle propagation has the
    reference's shapes and values."""
    N, S = 8, 8
    rng = np.random.default_rng(5)
    rec = m.scenarios.generate(1, 2, seed=57, weight_goal_1=20.0)
    obst = random_obstacles(rng, 1, 2, S, rec)
    ocfg = o2.default_config(2)
    for robot in (0, 1):
        pl, goal = P.set_planner_panda(7, 0, S, LINKS, {}, MOUNT, robot)
        r, o = rec[0, robot], obst[0, robot]
        fr = P.FabricsRollouts(N=N, dt=0.01, nx=7, nu=7, dof=7, nr_obsts=0, bool_ring=False, nr_obsts_dyn=S,
                               v_obsts_dyn=[o[i, 3:6] for i in range(S)], fabrics_mode="vel", collision_links_nrs=LINKS,
                               nr_constraints=1, constraints=np.array([0, 0, 1, -0.65]), nr_goals=3)
        fr.preset_radii_obsts_dyn([0.08] * S)
        fr.symbolic_forward_fabrics(pl, goal)
        x0 = [o[i, 0:3] for i in range(S)]
        x_goals = {"subgoal0": r[14:17], "subgoal1": np.array([0.107, 0, 0]), "subgoal2": np.array([np.pi / 4])}
        w_goals = {"subgoal0": r[17], "subgoal1": 20.0, "subgoal2": 1.0}
        q_st, qd_st, qdd_st = fr.forward_fabrics(planner=pl, pos_k=r[0:7], vel_k=r[7:14], ob_robot={}, goal=goal,
                                                 x_obsts_dyn_0=x0, x_goals_struct=x_goals, weight_goals_struct=w_goals)
        rq, rqd, _ = o2.rollout_cartesian(ocfg, robot, r, o[:, 0:3], o[:, 3:6], o[:, 9], N)
        assert len(q_st) == N and len(qd_st) == N and qdd_st == []
        assert np.abs(np.array(q_st) - rq).max() < 1e-9 and np.abs(np.array(qd_st) - rqd).max() < 1e-9
        # one action, and the symbolic twin of the same horizon
        a0 = fr.get_action(pl, r[0:7], r[7:14], x_obsts=[], x_obsts_dyn=x0, x_goals=list(x_goals.values()),
                           weight_goals=list(w_goals.values()))
        assert np.abs(a0 - rqd[0]).max() < 1e-9
        args = fr.define_arguments_numerical(q_robot=r[0:7], q_dot_robot=r[7:14], weight_goals=w_goals, x_goals=x_goals,
                                             x_obsts=[], x_obsts_dyn=x0, v_obsts_dyn=fr.v_obsts_dyn,
                                             constraints=np.array([0, 0, 1, -0.65]))
        q_n, _, _ = fr.rollouts_numerical(args)
        assert np.abs(q_n - np.array(q_st).T).max() < 1e-9
        # constant-velocity obstacle propagation
        x_N, x_list = fr.get_x_obsts_dyn_N(x0)
        assert len(x_N) == N + 1 and x_N[0].shape == (3, S) and len(x_list) == N
        for k in range(N):
            assert np.abs(x_N[k] - (o[:, 0:3] + k * 0.01 * o[:, 3:6]).T).max() < 1e-15
            assert np.abs(np.array(list(x_list[k])) - (o[:, 0:3] + k * 0.01 * o[:, 3:6])).max() < 1e-15
        xs = fr.x_obsts_dyn_numerical(x0)
        assert len(xs) == N and xs[0].shape == (3, S)
        assert np.abs(xs[N - 1] - (o[:, 0:3] + N * 0.01 * o[:, 3:6]).T).max() < 1e-15
        p1, v1 = fr.system_step(r[0:7], r[7:14], a0, dt=0.01, fabrics_mode="vel")
        assert np.array_equal(p1, r[0:7] + 0.01 * a0) and np.array_equal(v1, a0)


def test_jointspace_rollouts_numerical_obstacles(built):
    """ForwardFabricsPlanner.rollouts_numerical_obstacles (forward_planner_Jointspace.py:425-511): the other robots'
    sphere positions / velocities / accelerations along the horizon, against FK of the oracle's trajectories."""
    R, N = 3, 6
    rec = m.scenarios.generate(1, R, seed=58)
    params = types.SimpleNamespace(N_HORIZON=N, dt=0.01, dof=[7] * R, nr_obsts=[0] * R, fabrics_mode="vel",
                                   r_robots=[[0.08] * 8] * R, rotation_matrix_pandas=[P.ROT_PANDA] * R,
                                   collision_links_nrs=[LINKS] * R, STATIC_OR_DYN_FABRICS=1)
    mounts = {"mount_positions": [np.array([0.0, 0.0, 0.65]), np.array([1.0, 0.0, 0.65]), np.array([0.7, 0.6, 0.65])]}
    planners = [P.set_planner_panda(7, 0, 8 * (R - 1), LINKS, {}, mounts, i)[0] for i in range(R)]
    fwd = P.ForwardFabricsPlanner(params, planners, N_steps=N)
    ia = {"q_robots": [rec[0, i, 0:7] for i in range(R)], "q_dot_robots": [rec[0, i, 7:14] for i in range(R)],
          "x_obsts": [[] for _ in range(R)], "x_goals0": [rec[0, i, 14:17] for i in range(R)],
          "x_goals1": [rec[0, i, 18:21] for i in range(R)], "x_goals2": [rec[0, i, 22:23] for i in range(R)],
          "weight_goals0": [rec[0, i, 17] for i in range(R)], "weight_goals1": [rec[0, i, 21] for i in range(R)],
          "weight_goals2": [rec[0, i, 23] for i in range(R)], "constraints": [rec[0, i, 33:37] for i in range(R)]}
    xs, vs, as_ = fwd.rollouts_numerical_obstacles(ia)
    ocfg = o2.default_config(R)
    qN, qdN, _, _ = o2.rollout_jointspace(ocfg, rec[0], N)
    for i in range(R):
        others = [j for j in range(R) if j != i]
        assert len(xs[f"robot_{i}"]) == N and xs[f"robot_{i}"][0].shape == (3, 8 * (R - 1))
        for k in range(N):
            ex, ev, ea = [], [], []
            for j in others:
                qd_prev = rec[0, j, 7:14] if k == 0 else qdN[j, k - 1]
                x, v, c, _ = o2.kinematics(ocfg, j, qN[j, k], qd_prev)
                ex.append(x); ev.append(v); ea.append(ocfg.jdot_ref_sign * c)
            assert np.abs(xs[f"robot_{i}"][k] - np.concatenate(ex).T).max() < 1e-9
            assert np.abs(vs[f"robot_{i}"][k] - np.concatenate(ev).T).max() < 1e-9
            assert np.abs(as_[f"robot_{i}"][k] - np.concatenate(ea).T).max() < 1e-8


def test_deadlock_dropin_matches_reference_golden(built):
    """The CUDA deadlock kernel behind the reference's class interface replays the reference's own outputs bit for bit
    (goals, weights, time_deadlock_out) -- 'deadlock flags identical'."""
    g = np.load(os.path.join(GOLD, "deadlock_golden.npz"))
    raised = 0
    for c in range(int(g["n_cases"])):
        p = f"c{c}_"
        R = int(g[p + "R"])
        dl, ref = P.deadlockprevention([7] * R, R, 20), DeadlockOracle(R)
        tdo = tdo_r = 1000
        for t in range(len(g[p + "x"])):
            goals = [v.copy() for v in g[p + "goals"][t]]
            weights = [float(w) for w in g[p + "weights"][t]]
            go, wo, tdo = dl.deadlock_checking([v.copy() for v in g[p + "x"][t]], goals, weights, int(g[p + "time_step"][t]),
                                               tdo, float(g[p + "avg"][t]), list(g[p + "states"][t]))
            _, _, tdo_r, flag = ref.step(g[p + "x"][t], g[p + "goals"][t], g[p + "weights"][t], int(g[p + "time_step"][t]),
                                         tdo_r, float(g[p + "avg"][t]), list(g[p + "states"][t]))
            raised += int(dl.deadlock)
            assert dl.deadlock == flag
            assert np.array_equal(np.array(go), g[p + "goals_out"][t]), (c, t)
            assert np.array_equal(np.array(wo, dtype=float), g[p + "weights_out"][t]), (c, t)
            assert tdo == g[p + "tdo_out"][t]
            assert go is goals and wo is weights            # mutated in place like the reference
    assert raised > 50


def test_dropin_collision_link_subset_and_static_rollout_obstacles(built):
    """The reference's signature default collision_links_nr=[5] (example_pandas_Jointspace.py:64) through
    set_planner_panda / compute_action, and rollout planners with static obstacles (nr_obsts > 0,
    forward_planner_Jointspace.py:319-322) through ForwardFabricsPlanner -- both against the oracle."""
    rng = np.random.default_rng(12)
    rec = m.scenarios.generate(1, 2, seed=59, weight_goal_1=20.0)
    S = 4
    obst = random_obstacles(rng, 1, 2, S, rec)
    for links in ([5], [4, 7, 8]):
        pl, _ = P.set_planner_panda(7, 0, S, links, {}, MOUNT, 1)
        r, o = rec[0, 1], obst[0, 1]
        kw = dict(q=r[0:7], qdot=r[7:14], x_goal_0=r[14:17], weight_goal_0=r[17], angle_goal_1=P.ROT_PANDA,
                  x_goal_1=r[18:21], weight_goal_1=20.0, x_goal_2=r[22:23], weight_goal_2=1.0, x_obsts=[], radius_obsts=[],
                  constraint_0=r[33:37], radius_body_panda_links={str(l): np.array(0.08) for l in links},
                  radius_body_panda_hand=np.array(0.02), x_obsts_dynamic=[o[k, 0:3] for k in range(S)],
                  xdot_obsts_dynamic=[o[k, 3:6] for k in range(S)], xddot_obsts_dynamic=[o[k, 6:9] for k in range(S)],
                  radius_obsts_dynamic=[o[k, 9] for k in range(S)])
        act = pl.compute_action(**kw)
        ocfg = o2.set_collision_links(o2.default_config(2), [LINKS, links])
        rr = r.copy()
        rr[37:43] = [0.08 if l in links else 0.0 for l in range(3, 9)]
        ref = o2.action(ocfg, 1, rr, o[:, 0:3], o[:, 3:6], o[:, 6:9], o[:, 9])
        assert np.abs(act - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        with pytest.raises(KeyError):
            pl.compute_action(**dict(kw, radius_body_panda_links={}))
    # rollouts with two static spheres per robot
    R, N, Ss = 2, 8, 2
    rec = m.scenarios.generate(1, R, seed=60)
    params = make_params(R, N)
    params.nr_obsts = [Ss] * R
    params.radius_obsts = [[0.1, 0.07]] * R
    planners = [P.set_planner_panda(7, Ss, 8 * (R - 1), LINKS, {}, MOUNT, i)[0] for i in range(R)]
    fwd = P.ForwardFabricsPlanner(params, planners, N_steps=N)
    xs = rng.uniform([0.1, -0.5, 0.95], [0.9, 0.5, 1.4], size=(R, Ss, 3))
    ia = {"q_robots": [rec[0, i, 0:7] for i in range(R)], "q_dot_robots": [rec[0, i, 7:14] for i in range(R)],
          "x_obsts": [[xs[i, o] for o in range(Ss)] for i in range(R)], "x_goals0": [rec[0, i, 14:17] for i in range(R)],
          "x_goals1": [rec[0, i, 18:21] for i in range(R)], "x_goals2": [rec[0, i, 22:23] for i in range(R)],
          "weight_goals0": [rec[0, i, 17] for i in range(R)], "weight_goals1": [rec[0, i, 21] for i in range(R)],
          "weight_goals2": [rec[0, i, 23] for i in range(R)], "constraints": [rec[0, i, 33:37] for i in range(R)]}
    qN, qdN, avg, _ = o2.rollout_jointspace_static(o2.default_config(R), rec[0], N, xs, np.array(params.radius_obsts))
    va = fwd.get_velocity_rollouts(ia)
    assert np.abs(np.array([v[0] for v in va]) - avg).max() < 1e-9
    q_n, qd_n, _ = fwd.rollouts_numerical(ia)
    for i in range(R):
        assert np.abs(qd_n[f"robot_{i}"][0] - qdN[i].T).max() < 1e-9 and np.abs(q_n[f"robot_{i}"][0] - qN[i].T).max() < 1e-9


def test_deadlock_dropin_point_mass_branch(built):
    """deadlockprevention with dof[0] == 2 (deadlock_prevention.py:12-19): the kernel with the point-mass constants
    replays sequences produced by the reference's own class bit for bit; planar (2-D) positions behave like the
    reference too: they work until a deadlock fires, then the reference's goal_robot0[2] access raises IndexError."""
    g = np.load(os.path.join(GOLD, "deadlock_point_golden.npz"))
    raised = 0
    for c in range(int(g["n_cases"])):
        p = f"c{c}_"
        R = int(g[p + "R"])
        dl = P.deadlockprevention([2] * R, R, 20)
        assert (dl.avg_vel_constant, dl.dist_constant, dl.goal_weight_follower, dl.goal_weight_leader, dl.time_wait,
                dl.nr_goal_scale) == (0.03, 1.0, 10, 1, 50, 100)
        tdo = 1000
        for t in range(len(g[p + "x"])):
            goals = [v.copy() for v in g[p + "goals"][t]]
            weights = [float(w) for w in g[p + "weights"][t]]
            go, wo, tdo = dl.deadlock_checking([v.copy() for v in g[p + "x"][t]], goals, weights, int(g[p + "time_step"][t]),
                                               tdo, float(g[p + "avg"][t]), list(g[p + "states"][t]))
            raised += int(dl.deadlock)
            assert np.array_equal(np.array(go), g[p + "goals_out"][t]), (c, t)
            assert np.array_equal(np.array(wo, dtype=float), g[p + "weights_out"][t]), (c, t)
            assert tdo == g[p + "tdo_out"][t]
    assert raised > 50
    dl = P.deadlockprevention([2, 2], 2, 20)
    x2 = [np.array([0.0, 0.0]), np.array([0.1, 0.0])]
    go, wo, tdo = dl.deadlock_checking(x2, [np.array([2.0, 0.0]), np.array([-2.0, 0.0])], [1.0, 1.0], 5, 1000, 0.01, [0, 0])
    assert tdo == 1000 and np.asarray(go[0]).shape == (2,)                       # time gate closed: nothing happens
    with pytest.raises(IndexError):
        dl.deadlock_checking(x2, [np.array([2.0, 0.0]), np.array([-2.0, 0.0])], [1.0, 1.0], 50, 1000, 0.01, [0, 0])
    with pytest.raises(IndexError):
        dl.deadlock_checking(x2, [np.array([2.0, 0.0])], [1.0, 1.0], 50, 1000, 0.01, [0, 0])   # short list


def test_batched_deadlock_from_rollout_flags_identical(built):
    """FP64 rollout -> batched deadlock kernel with engineered end-effector positions (many candidate pairs): flags,
    follower goals and weights equal the oracle's on every scenario the oracle can evaluate -- no knife-edge mask."""
    import torch
    from multi_robot_fabrics_b200.api import to_soa
    R, N, B = 3, 20, 512
    rec = m.scenarios.generate(B, R, seed=54)
    rec[:, :, 7:14] *= 0.1                                  # slow scenarios so that avg_vel < 0.16 happens
    fab = Fabrics(R, device=0, estimate_goal=1)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N, estimate_goal=1)
    states = np.random.default_rng(1).choice([0, 1, 1, 0, 2], size=(B, R)).astype(np.int32)
    exp_flag = np.zeros(B, dtype=np.int32)
    exp_goals = rec[:, :, 14:17].copy()
    exp_goals[:, 1] = goal
    exp_w = rec[:, :, 17].copy()
    xe = xee.copy()
    xe[::2, 1] = xe[::2, 0] + [0.05, 0.02, 0.01]            # make every other scenario a near-contact pair (0,1)
    for b in range(B):
        d = DeadlockOracle(R)
        go, wo, _, fl = d.step(xe[b], exp_goals[b], exp_w[b], 100, 1000, float(sum(avg[b]) / R), list(states[b]))
        exp_flag[b], exp_goals[b], exp_w[b] = fl, np.array(go), np.array(wo, dtype=float)
    dt, dev = torch.float64, "cuda:0"
    d_rec = torch.from_numpy(to_soa(rec)).to(dev, dtype=dt)
    a = torch.empty((R, B), dtype=dt, device=dev)
    x = torch.empty((R, 3, B), dtype=dt, device=dev)
    ge = torch.empty((3, B), dtype=dt, device=dev)
    fab.rollout_dev(d_rec, N, avg_vel=a, x_ee=x, goal_est=ge)
    x = torch.from_numpy(np.ascontiguousarray(xe.transpose(1, 2, 0))).to(dev, dtype=dt)   # the engineered ee positions
    goals = d_rec[14:17].permute(1, 0, 2).contiguous()
    goals[1] = ge
    w = d_rec[17].clone()
    sm = torch.from_numpy(np.ascontiguousarray(states.T)).to(dev)
    ts = torch.full((B,), 100, dtype=torch.int32, device=dev)
    tdo = torch.full((B,), 1000, dtype=torch.int32, device=dev)
    st_int = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=dev).repeat_interleave(B).contiguous()
    st_goal = torch.zeros((3, B), dtype=dt, device=dev)
    flag = fab.deadlock_dev(x, goals, w, sm, ts, tdo, st_int, st_goal, avg_vel=a)
    torch.cuda.synchronize()
    assert np.array_equal(flag.cpu().numpy()[ok], exp_flag[ok])
    assert exp_flag[ok].sum() > 20
    got = goals.permute(2, 0, 1).double().cpu().numpy()
    assert np.abs(got - exp_goals)[ok].max() < 1e-12
    assert np.array_equal(w.T.double().cpu().numpy()[ok], exp_w[ok])
    fab.close()


@pytest.mark.parametrize("R,N,B,seed", [(2, 20, 4096, 61), (3, 50, 2048, 62), (3, 20, 8192, 63)])
def test_fused_rfcv_step_deadlock_flags_identical(built, R, N, B, seed):
    """'Deadlock flags identical' for the path bench.py times (BASELINE configs C3 and C5 shapes, and the metric shape):
    rollout kernel -> mrf_rfcv_post_dev.  For FP32 the scenarios whose results sit in the guard band of a threshold, are
    numerically stiff or non-finite are re-rolled by the FP64 kernel inside the post step, so the flags -- and the
    resolved goals / weights -- equal the float64 oracle's on EVERY scenario the oracle can evaluate, no mask.  Both
    sides consume the same (float32-representable) records.  The end-effector distance test (0.35 m in the reference) is
    widened through MrfConfig so that random scenarios raise flags."""
    import torch
    from multi_robot_fabrics_b200.api import to_soa
    DIST = 1.0
    rec = m.scenarios.generate(B, R, seed=seed).astype(np.float32).astype(np.float64)
    rec[:, :, 7:14] = (rec[:, :, 7:14] * 0.25).astype(np.float32)      # vel_avg_tot spread around the 0.16 threshold
    ref = o2.rollout_rfcv(o2.default_config(R), rec, N)
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(ref["avg_vel"]).all(axis=1) & (ref["avg_vel"].max(axis=1) < 3.0 ** 2)
    states = np.random.default_rng(seed).choice([0, 1, 1, 0, 0, 2], size=(B, R)).astype(np.int32)
    exp_flag, exp_tdo = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    exp_goals = rec[:, :, 14:17].copy()
    exp_goals[:, 1] = ref["goal_est"]
    exp_w = rec[:, :, 17].copy()
    for b in np.nonzero(ok)[0]:
        go, wo, exp_tdo[b], fl = DeadlockOracle(R, dist_endeff=DIST).step(
            ref["x_ee"][b], exp_goals[b], exp_w[b], 100, 1000, float(sum(ref["avg_vel"][b]) / R), list(states[b]))
        exp_flag[b], exp_goals[b], exp_w[b] = fl, np.array(go), np.array(wo, dtype=float)
    near = np.abs(ref["avg_vel"].sum(axis=1) / R - 0.16)[ok]
    assert exp_flag[ok].sum() > 20 and (near < 2e-3).sum() > 3           # the knife edge is populated
    fab = Fabrics(R, device=0, estimate_goal=1, dl_dist_endeff=DIST)
    dev = "cuda:0"
    for dt, gtol in ((torch.float64, 1e-12), (torch.float32, 2e-6)):
        d_rec = torch.from_numpy(to_soa(rec)).to(dev, dtype=dt)
        work = d_rec.clone()
        t = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
        a, x, ge, risk, result = t(R, B), t(R, 3, B), t(3, B), t(R, B), t(R + 1, B)
        use_risk = dt == torch.float32
        fab.rollout_dev(d_rec, N, avg_vel=a, x_ee=x, goal_est=ge, risk=risk if use_risk else None)
        sm = torch.from_numpy(np.ascontiguousarray(states.T)).to(dev)
        ts = torch.full((B,), 100, dtype=torch.int32, device=dev)
        tdo = torch.full((B,), 1000, dtype=torch.int32, device=dev)
        st_int = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=dev).repeat_interleave(B).contiguous()
        st_goal = torch.zeros((3, B), dtype=dt, device=dev)
        flag = fab.rfcv_post_dev(d_rec, N, x, work, ge, a, sm, ts, tdo, st_int, st_goal, risk=risk if use_risk else None,
                                 result=result)
        torch.cuda.synchronize()
        got_flag = flag.cpu().numpy()
        assert np.array_equal(got_flag[ok], exp_flag[ok]), (str(dt), int((got_flag[ok] != exp_flag[ok]).sum()))
        assert np.array_equal(tdo.cpu().numpy()[ok], exp_tdo[ok])
        assert np.array_equal(result[R].cpu().numpy()[ok], exp_flag[ok].astype(np.float64))
        got_goals = work[14:17].permute(2, 1, 0).double().cpu().numpy()
        assert np.abs(got_goals - exp_goals)[ok].max() < gtol
        assert np.array_equal(work[17].T.double().cpu().numpy()[ok], exp_w[ok])
        assert np.abs(result[:R].T.double().cpu().numpy() - ref["avg_vel"])[ok].max() < (1e-9 if dt == torch.float64 else 5e-2)
        if use_risk:
            rerolled, overflow, listed = fab.guard_stats()
            assert overflow == 0 and 0 < listed < B // 8, (rerolled, overflow, listed)
    fab.close()


def test_rfcv_post_step_repeats_and_overflows_cleanly(built):
    """The post step's list counter is reset by the last CTA of the FP64 re-roll kernel: calling the step again gives the
    same list and the same flags; a list that overflows its capacity re-rolls `cap` scenarios, decides the others from the
    FP32 values, and leaves the counters clean for the next call."""
    import torch
    from multi_robot_fabrics_b200.api import to_soa
    R, N, B = 3, 20, 8192
    rec = m.scenarios.generate(B, R, seed=64).astype(np.float32)
    rec[:, :, 7:14] *= 0.25
    fab = Fabrics(R, device=0, estimate_goal=1, dl_dist_endeff=1.0)
    dev = "cuda:0"
    d_rec = torch.from_numpy(to_soa(rec)).to(dev)
    t = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    a, x, ge, risk = t(R, B), t(R, 3, B), t(3, B), t(R, B)
    fab.rollout_dev(d_rec, N, avg_vel=a, x_ee=x, goal_est=ge, risk=risk)
    sm = torch.zeros((R, B), dtype=torch.int32, device=dev)
    ts = torch.full((B,), 100, dtype=torch.int32, device=dev)

    def post():
        work, result = d_rec.clone(), t(R + 1, B)
        tdo = torch.full((B,), 1000, dtype=torch.int32, device=dev)
        st_int = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=dev).repeat_interleave(B).contiguous()
        flag = fab.rfcv_post_dev(d_rec, N, x, work, ge, a, sm, ts, tdo, st_int, t(3, B).zero_(), risk=risk, result=result)
        torch.cuda.synchronize()
        return flag.cpu().numpy(), result.cpu().numpy(), fab.guard_stats()

    f1, r1, (rr1, ov1, listed1) = post()
    f2, r2, (rr2, ov2, listed2) = post()
    assert listed1 > 40 and listed2 == listed1 and ov1 == ov2 == 0 and rr2 == 2 * rr1
    assert np.array_equal(f1, f2) and np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
    fab.set_guard(cap=32)
    f3, r3, (rr3, ov3, listed3) = post()
    assert listed3 == listed1 and ov3 == listed1 - 32 and rr3 == rr2 + 32
    assert (f3 != f1).sum() <= listed1 - 32          # only scenarios that lost their re-roll may differ
    fab.set_guard(cap=0)                             # back to the default capacity
    f4, r4, (rr4, ov4, listed4) = post()
    assert listed4 == listed1 and ov4 == ov3 and np.array_equal(f4, f1)
    fab.close()


def test_rfcv_host_submit_end_to_end_matches_device_path(built):
    """mrf_rfcv_host_submit_f32 -- the whole RF-CV step from page-locked host records (in-place rollout kernel, FP64 guard
    re-roll reading the listed scenarios from the same host records, deadlock heuristic, result back to host) -- returns
    bit for bit what the device-tensor path (rollout_dev + rfcv_post_dev) returns, with whole and with compact records."""
    import torch
    from multi_robot_fabrics_b200.api import to_soa
    R, N, B = 3, 20, 4099                                   # ragged last tile
    rec = m.scenarios.generate(B, R, seed=64).astype(np.float32)
    rec[:, :, 7:14] *= np.float32(0.25)
    fab = Fabrics(R, device=0, estimate_goal=1, dl_dist_endeff=1.0)
    dev = "cuda:0"
    d_rec = torch.from_numpy(to_soa(rec)).to(dev)
    work = d_rec.clone()
    t = lambda *shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)
    a, x, ge, risk, res = t(R, B), t(R, 3, B), t(3, B), t(R, B), t(R + 1, B)
    fab.rollout_dev(d_rec, N, avg_vel=a, x_ee=x, goal_est=ge, risk=risk)
    fab.rfcv_post_dev(d_rec, N, x, work, ge, a, torch.zeros((R, B), dtype=torch.int32, device=dev),
                      torch.full((B,), 100, dtype=torch.int32, device=dev), torch.full((B,), 1000, dtype=torch.int32, device=dev),
                      torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=dev).repeat_interleave(B).contiguous(),
                      torch.zeros((3, B), device=dev), risk=risk, result=res)
    torch.cuda.synchronize()
    want = res.cpu().numpy()
    want_gw = torch.cat([work[14:17], work[17:18]]).cpu().numpy()                       # (4,R,B) after the heuristic
    listed = fab.guard_stats()[2]
    assert want[R].sum() > 20 and listed > 0
    pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).pin_memory().numpy()
    for compact in (False, True):
        h_rec = pin(rec[:, :, :18] if compact else rec)
        h_res, h_gw = pin(np.zeros((R + 1, B))), pin(np.zeros((4, R, B)))
        for _ in range(3):                                                               # both pipeline slots
            fab.rfcv_host_submit(h_rec, N, h_res, shared=rec[0] if compact else None, time_step=100, goals_out=h_gw)
        fab.rollout_host_wait(all=True)
        assert np.array_equal(h_res.view(np.uint32), want.view(np.uint32)), compact
        assert np.array_equal(h_gw.view(np.uint32), want_gw.view(np.uint32)), compact
    with pytest.raises(m.MrfError):
        fab.rfcv_host_submit(np.ascontiguousarray(rec), N, h_res)                      # pageable records are refused
    fab.close()


def test_point_mass_planner_config_c1(built):
    """BASELINE config C1: the 4-point-mass examples through the drop-in planner, against the O1 golden actions and
    (batched, random states) against oracle O2."""
    import torch
    g = np.load(os.path.join(GOLD, "fabric_golden.npz"))
    pos, vel, goals, obst = g["pm_pos"], g["pm_vel"], g["pm_goals"], g["pm_obst"]
    pl_s = P.set_planner_point(None, n_obstacles=9)
    pl_d = P.set_planner_point(None, n_obstacles=6, n_dyn_obstacles=3)
    for i in range(4):
        others = [j for j in range(4) if j != i]
        a = pl_s.compute_action(q=pos[i], qdot=vel[i], x_goal_0=goals[i], weight_goal_0=1.0,
                                x_obsts=list(obst) + [pos[j] for j in others], radius_obsts=[1.0] * 6 + [0.2] * 3,
                                radius_body_base_link=np.array(0.2))
        assert np.abs(a - g["pm_static"][i]).max() < 1e-10
        kw = dict(q=pos[i], qdot=vel[i], x_goal_0=goals[i], weight_goal_0=1.0, x_obsts=list(obst), radius_obsts=[1.0] * 6,
                  radius_body_base_link=np.array(0.2))
        for k, j in enumerate(others):                       # per-index spellings, example_pointmasses_dynamic.py:199-211
            kw[f"x_obst_dynamic_{k}"], kw[f"xdot_obst_dynamic_{k}"] = pos[j][0:2], vel[j][0:2]
            kw[f"xddot_obst_dynamic_{k}"], kw[f"radius_obst_dynamic_{k}"] = np.array([0, 0]), np.array(0.2)
        a = pl_d.compute_action(**kw)
        assert np.abs(a - g["pm_dyn"][i]).max() < 1e-10
    # batched device entry vs O2
    rng = np.random.default_rng(8)
    B, Ss, Sd = 300, 6, 3
    rec = np.zeros((B, 10))
    rec[:, 0:2] = rng.uniform(-3, 3, (B, 2))
    rec[:, 3:6] = rng.uniform(-1, 1, (B, 3))
    rec[:, 6:8] = rng.uniform(-3, 3, (B, 2))
    rec[:, 8], rec[:, 9] = 1.0, 0.2
    stat = np.zeros((B, Ss, 4))
    ang = rng.uniform(0, 2 * np.pi, (B, Ss))
    rad = rng.uniform(2.0, 4.0, (B, Ss))                    # keep clear of the robot: normalised clearance >= 0.6
    stat[:, :, 0] = rec[:, None, 0] + rad * np.cos(ang)
    stat[:, :, 1] = rec[:, None, 1] + rad * np.sin(ang)
    stat[:, :, 3] = 1.0
    dyn = np.zeros((B, Sd, 7))
    ang = rng.uniform(0, 2 * np.pi, (B, Sd))
    dyn[:, :, 0] = rec[:, None, 0] + 1.5 * np.cos(ang)
    dyn[:, :, 1] = rec[:, None, 1] + 1.5 * np.sin(ang)
    dyn[:, :, 2:6] = rng.uniform(-0.5, 0.5, (B, Sd, 4))
    dyn[:, :, 6] = 0.2
    fab = pl_s.fab
    for dt, tol in ((torch.float64, 1e-9), (torch.float32, 2e-3)):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.moveaxis(a, 0, -1))).to("cuda:0", dtype=dt)
        act = fab.point_action_dev(t(rec), t(stat), t(dyn))
        torch.cuda.synchronize()
        got = act.T.double().cpu().numpy()
        ocfg = o2.default_config(2)
        for b in range(B):
            ref = o2.point_action(ocfg, rec[b, 0:3], rec[b, 3:6], rec[b, 6:8], 1.0, 0.2, stat[b, :, 0:3], stat[b, :, 3],
                                  dyn[b, :, 0:2], dyn[b, :, 2:4], dyn[b, :, 4:6], dyn[b, :, 6])
            assert np.abs(got[b] - ref).max() < tol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("R,kw", [(2, dict(rollout_fabrics=True, resolve_deadlocks=True, estimate_goal=True)),
                                  (3, dict(rollout_fabrics=True, resolve_deadlocks=True, estimate_goal=False)),
                                  (2, dict(rollout_fabrics=False))])
def test_closed_loop_episodes_match_cpu_loop(built, R, kw):
    """the batched closed control loop (rollout -> deadlock -> obstacle staging -> action -> kinematic
    step, captured in a CUDA graph) reproduces the same loop run on the CPU oracle, scenario by scenario."""
    from multi_robot_fabrics_b200.episodes import BatchedEpisodes
    B, T, N = 12, 25, 5
    rec = m.scenarios.generate(B, R, seed=71)
    rec[:, :, 7:14] *= 0.2
    rec[1::2, 1, 14:17] = rec[1::2, 0, 14:17] + [0.05, 0.0, 0.02]      # half of the scenarios: nearly the same goal
    ep = BatchedEpisodes(rec, n_horizon=N, dtype="f64", n_obst_per_link=2, **kw).run(T)
    res = ep.results()
    n_cmp = 0
    for b in range(B):
        try:
            q, n_flags, done_at = oracle_episode(rec[b], T, N, rollout=kw.get("rollout_fabrics", True),
                                                 resolve=kw.get("resolve_deadlocks", True),
                                                 estimate=kw.get("estimate_goal", False), n_per_link=2)
        except FloatingPointError:
            continue
        if not np.isfinite(q).all():
            continue
        n_cmp += 1
        assert np.abs(res["q"][b] - q).max() < 1e-7, b
        assert res["deadlock_steps"][b] == n_flags and res["steps_to_success"][b] == done_at
        assert res["nonfinite_steps"][b] == 0          # an episode the oracle follows has no non-finite action
    assert n_cmp >= B - 3 and res["steps"] == T
    # a scenario started beyond a joint limit loses positive definiteness (NaN action): the arm holds still and the step
    # is COUNTED instead of silently vanishing from the metrics
    bad = rec[:2].copy()
    bad[0, 0, 3], bad[0, 0, 10] = 0.5, 0.3               # joint 4 above its upper limit (-0.0698), moving further out
    rb = BatchedEpisodes(bad, n_horizon=N, dtype="f64", n_obst_per_link=2, use_graph=False, **kw).run(3).results()
    assert rb["nonfinite_steps"][0] >= 1


def test_deadlock_in_place_on_record_tensor(built):
    """mrf_deadlock_rec_dev (goals / weights rows of the SoA record tensor, RF-CV estimate applied first) equals the
    separate-array entry."""
    import torch
    from multi_robot_fabrics_b200.api import to_soa
    R, N, B = 3, 10, 300
    rec = m.scenarios.generate(B, R, seed=55)
    rec[:, :, 7:14] *= 0.1
    fab = Fabrics(R, device=0, estimate_goal=1)
    dev = "cuda:0"
    for dt in (torch.float64, torch.float32):
        d_rec = torch.from_numpy(to_soa(rec)).to(dev, dtype=dt)
        a = torch.empty((R, B), dtype=dt, device=dev)
        x = torch.empty((R, 3, B), dtype=dt, device=dev)
        ge = torch.empty((3, B), dtype=dt, device=dev)
        fab.rollout_dev(d_rec, N, avg_vel=a, x_ee=x, goal_est=ge)
        x[1, :, ::2] = x[0, :, ::2] + 0.03                     # near-contact hands in every other scenario
        sm = torch.zeros((R, B), dtype=torch.int32, device=dev)
        ts = torch.full((B,), 50, dtype=torch.int32, device=dev)
        mk = lambda: (torch.full((B,), 1000, dtype=torch.int32, device=dev),
                      torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=dev).repeat_interleave(B).contiguous(),
                      torch.zeros((3, B), dtype=dt, device=dev))
        goals = d_rec[14:17].permute(1, 0, 2).contiguous()
        goals[1] = ge
        w = d_rec[17].clone()
        tdo1, si1, sg1 = mk()
        f1 = fab.deadlock_dev(x, goals, w, sm, ts, tdo1, si1, sg1, avg_vel=a)
        work = d_rec.clone()
        tdo2, si2, sg2 = mk()
        f2 = fab.deadlock_rec_dev(x, work, sm, ts, tdo2, si2, sg2, goal_est=ge, avg_vel=a)
        torch.cuda.synchronize()
        assert torch.equal(f1, f2) and f1.sum() > 20 and torch.equal(tdo1, tdo2) and torch.equal(si1, si2)
        assert torch.equal(work[14:17].permute(1, 0, 2), goals) and torch.equal(work[17], w)
        assert torch.equal(work[0:14], d_rec[0:14]) and torch.equal(work[18:], d_rec[18:])
    fab.close()


def test_fsm_kernel_matches_reference_golden(built):
    """The batched state-machine kernel replays the reference class's own sequences (all 8 cases as one batch): state,
    goal, weight and gripper action identical at every step; plus the StateMachine drop-in class on one case."""
    import torch
    g = np.load(os.path.join(GOLD, "fsm_golden.npz"))
    C_ = int(g["n_cases"])
    T = len(g["c0_x"])
    dev = "cuda:0"
    for nb in sorted(set(int(g[f"c{c}_nr_blocks"]) for c in range(C_))):
        cases = [c for c in range(C_) if int(g[f"c{c}_nr_blocks"]) == nb]
        B = len(cases)
        fab = Fabrics(1, device=0)
        col = lambda key, t: torch.tensor(np.stack([g[f"c{c}_{key}"][t] for c in cases], axis=-1)[None], dtype=torch.float64, device=dev)
        start = torch.tensor(np.stack([g[f"c{c}_start"] for c in cases], axis=-1)[None], dtype=torch.float64, device=dev)
        goal, above = start.clone(), torch.zeros_like(start)
        weight = torch.full((1, B), 2.0, dtype=torch.float64, device=dev)
        st = torch.zeros((6, 1, B), dtype=torch.int32, device=dev)
        st[0] = 1
        grip = torch.zeros((1, 2, B), dtype=torch.float64, device=dev)
        for t in range(T):
            fab.fsm_dev([nb], col("x", t), col("qg", t), col("gb", t), start, goal, above, weight, st, grip)
            exp_state = np.array([g[f"c{c}_state"][t] for c in cases])
            assert np.array_equal(st[0, 0].cpu().numpy(), exp_state), t
            assert np.array_equal(goal[0].cpu().numpy().T, np.stack([g[f"c{c}_goal"][t] for c in cases]))
            assert np.array_equal(weight[0].cpu().numpy(), np.array([g[f"c{c}_weight"][t] for c in cases]))
            assert np.array_equal(grip[0].cpu().numpy().T, np.stack([g[f"c{c}_grip"][t] for c in cases]))
        fab.close()
    # drop-in class
    holder = {}
    sm = P.StateMachine(g["c0_start"], 2, int(g["c0_nr_blocks"]), lambda q: holder["x"], ["panda", "panda"])
    for t in range(0, 300):
        holder["x"] = g["c0_x"][t]
        s = sm.get_state_machine_panda(np.zeros(7), g["c0_qg"][t], g["c0_gb"][t], "panda")
        assert s == g["c0_state"][t] and np.array_equal(sm.get_goal_robot(), g["c0_goal"][t])
        assert sm.get_weight_goal0() == g["c0_weight"][t]
        assert np.array_equal(sm.get_gripper_action_panda(g["c0_qg"][t]), g["c0_grip"][t])


def test_utils_kinematics_dropins(built):
    """UtilsKinematics / utils_apply_fk drop-ins used the way example_pandas_Jointspace.py:324-343 and
    example_pandas_cartesian.py:361-418 use the CasADi functions."""
    from multi_robot_fabrics_b200 import utils as U
    R = 2
    rng = np.random.default_rng(5)
    planners = [P.set_planner_panda(7, 0, 8 * (R - 1), LINKS, "panda", MOUNT, i)[0] for i in range(R)]
    cfg = o2.default_config(R, jdot_ref_sign=-1.0)
    links = [[f"panda_link{k + 1}" for k in range(8)]] * R
    uk = U.UtilsKinematics()
    fk = uk.define_forward_kinematics(planners, [LINKS] * R, links)
    ee = uk.define_symbolic_endeffector(planners)
    q = rng.uniform(-1, 1, (R, 7)) + np.array([0, 0, 0, -1.5, 0, 1.8, 0])
    qd = rng.uniform(-1, 1, (R, 7))
    for i in range(R):
        x, v, c, J = o2.kinematics(cfg, i, q[i], qd[i])
        for z in range(8):
            np.testing.assert_allclose(fk["fk_fun"][i][z](q[i]).full().transpose()[0], x[z], atol=1e-12)
            jac = fk["jac_fun"][i][z](q[i])
            np.testing.assert_allclose(jac.full(), J[z], atol=1e-12)
            np.testing.assert_allclose((jac @ qd[i]).full().transpose()[0], v[z], atol=1e-12)
            # utils.py:28,37: Jdot_sign (-1) times d(J qd)/dq qd
            np.testing.assert_allclose((fk["jac_dot_fun"][i][z](q[i], qd[i]) @ qd[i]).full().transpose()[0], -c[z],
                                       atol=1e-11)
    x_ee, v_ee = U.compute_endeffector(list(q), list(qd), ee, nr_robots=R)
    for i in range(R):
        xo, vo = o2.endeffector(cfg, i, q[i], qd[i], use_jqd=True)
        np.testing.assert_allclose(x_ee[i], xo, atol=1e-12)
        np.testing.assert_allclose(v_ee[i], vo, atol=1e-12)
    # collision-sphere functions + compute_x_obsts_dyn_0 (n_obst_per_link = 2)
    n = 2
    off = o2.sphere_offsets_ref(n)
    T = [[[np.block([[np.eye(3), off[l, s].reshape(3, 1)], [np.zeros((1, 3)), np.ones((1, 1))]]) for s in range(n)]
          for l in range(8)] for _ in range(R)]
    mounts = [planners[i].mount for i in range(R)]
    sph = uk.define_symbolic_collision_link_poses({"URDF_file_panda": "unused"}, links, T, n_obst_per_link=n,
                                                  mount_transform=mounts)
    env = {}
    for i in range(R):
        xs, _, vs = o2.spheres(cfg, i, q[i], qd[i], off)
        np.testing.assert_allclose(sph[i]["fk_fun"](np.append(q[i], 0)).full().T, xs, atol=1e-12)
        np.testing.assert_allclose(sph[i]["vel_fun"](np.append(q[i], 0), np.append(qd[i], 0)).full().T, vs, atol=1e-12)
        for k in range(8 * n):
            env[(f"robot_{i}", k)] = xs[k]
    xd, vd, per = U.compute_x_obsts_dyn_0(list(q), list(qd), x_collision_sphere_poses=env, nr_robots=R,
                                          fk_dict_spheres=sph, nr_dyn_obsts=[8 * n] * R)
    ref = o2.obstacle_lists(cfg, q, qd, off, vel_mode=1)
    for i in range(R):
        np.testing.assert_allclose(np.array(xd[i]), ref[i][:, 0:3], atol=1e-12)
        np.testing.assert_allclose(np.array(vd[i]).reshape(-1, 3), ref[i][:, 3:6], atol=1e-12)


@pytest.mark.parametrize("kw", [dict(rollout_fabrics=True, resolve_deadlocks=True, estimate_goal=True),
                                dict(rollout_fabrics=False)])
def test_pick_and_place_episodes_match_cpu_loop(built, kw):
    """SURVEY 8f ranks 2 + 3 together: the reference's pick-and-place protocol (state machine -> goals / weights / planner
    choice / gripper, examples/example_pandas_Jointspace.py:289-312,417-448) inside the fused device control step,
    against the same loop on the CPU with oracle O2 and the state-machine / deadlock restatements pinned to the
    reference's classes.  Blocks sit a few centimetres from the hands so every state is visited within the test."""
    from helpers import oracle_pick_and_place
    from multi_robot_fabrics_b200.episodes import BatchedEpisodes
    R, B, T, N, nb = 2, 4, 420, 3, 1
    rng = np.random.default_rng(9)
    rec = m.scenarios.generate(B, R, seed=91)
    rec[:, :, 7:14] = 0.0
    cfg = o2.default_config(R)
    start = np.zeros((B, R, 3))
    blocks = np.zeros((B, nb, R, 3))
    for b in range(B):
        for r in range(R):
            start[b, r] = o2.kinematics(cfg, r, rec[b, r, 0:7], rec[b, r, 7:14])[0][7]
            blocks[b, 0, r] = start[b, r] + np.array([rng.uniform(-0.06, 0.06), rng.uniform(-0.06, 0.06), -0.16])
    ep = BatchedEpisodes(rec, n_horizon=N, dtype="f64", n_obst_per_link=1, blocks=blocks, start_goal=start, **kw).run(T)
    res = ep.results()
    for b in range(B):
        q, states, picked, n_flags, done_at, q_grip = oracle_pick_and_place(
            rec[b], blocks[b], start[b], T, N, rollout=kw.get("rollout_fabrics", True), resolve=kw.get("resolve_deadlocks", True),
            estimate=kw.get("estimate_goal", False))
        assert np.array_equal(res["state"][b], states) and np.array_equal(res["blocks_picked"][b], picked), b
        assert res["steps_to_success"][b] == done_at and res["deadlock_steps"][b] == n_flags
        assert np.abs(res["q"][b] - q).max() < 1e-6 and np.abs(res["q_grip"][b] - q_grip).max() < 1e-12
    assert res["blocks_picked"].sum() > 0          # the protocol really ran through grasp and release
