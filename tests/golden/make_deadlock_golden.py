"""Generates tests/golden/deadlock_golden.npz by running the REFERENCE's own deadlockprevention class
(imported from /root/reference; pure numpy) on seeded input sequences.  Run in the build container only:
    python tests/golden/make_deadlock_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/multi_robot_fabrics/others_planner/deadlock_prevention.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_deadlock_prevention", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.deadlockprevention


def make_sequence(rng, R, T):
    """A control-loop-like sequence: end effectors drift towards each other, velocities decay, states vary."""
    x = rng.uniform([0.2, -0.3, 0.9], [0.8, 0.3, 1.3], size=(R, 3))
    seq = []
    for t in range(T):
        x = x + rng.normal(0, 0.01, size=(R, 3)) + 0.02 * (x.mean(axis=0) - x)
        goals = rng.uniform([0.2, -0.6, 0.8], [0.8, 0.6, 1.25], size=(R, 3))
        weights = rng.choice([2.0, 3.0], size=R)
        avg = float(rng.choice([0.01, 0.05, 0.15, 0.159999, 0.16, 0.17, 0.5]) * rng.uniform(0.9, 1.1))
        states = rng.choice([0, 1, 0, 1, 0, 1, 2, 3], size=R)
        seq.append((x.copy(), goals, weights, t if rng.random() > 0.1 else int(rng.integers(0, 12)), avg, states))
    return seq


def main():
    DL = load_reference()
    rng = np.random.default_rng(2024)
    out = {}
    case = 0
    for R in (2, 3):
        for rep in range(6):
            T = 60
            dl = DL([7] * R, R, 20)
            tdo = 1000                                    # example_pandas_Jointspace.py:275
            seq = make_sequence(rng, R, T)
            X, G, W, TS, AV, ST = [], [], [], [], [], []
            GO, WO, TO = [], [], []
            for (x, goals, weights, ts, avg, states) in seq:
                gl = [g.copy() for g in goals]
                wl = [float(w) for w in weights]
                X.append(x); G.append(goals.copy()); W.append(weights.copy()); TS.append(ts); AV.append(avg); ST.append(states)
                g_o, w_o, tdo = dl.deadlock_checking(x_robots=[xi.copy() for xi in x], goal_robots=gl, goal_weights=wl,
                                                     time_step=ts, time_deadlock_out=tdo, avg_sum=avg,
                                                     state_machine_robots=list(states))
                GO.append(np.array([np.asarray(g, dtype=np.float64) for g in g_o]))
                WO.append(np.array([float(w) for w in w_o]))
                TO.append(tdo)
            pre = f"c{case}_"
            out[pre + "R"] = np.array(R)
            for k, v in (("x", X), ("goals", G), ("weights", W), ("time_step", TS), ("avg", AV), ("states", ST),
                         ("goals_out", GO), ("weights_out", WO), ("tdo_out", TO)):
                out[pre + k] = np.array(v)
            case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "deadlock_golden.npz"), **out)
    print("wrote", case, "cases")
    # ---- point-mass branch of the class (dof[0] == 2: deadlock_prevention.py:12-19), 3-D positions (the reference
    # indexes goal_robot0[2], :99, so planar vectors raise IndexError as soon as a deadlock fires) ----
    out, case = {}, 0
    for R in (2, 4):
        for rep in range(3):
            T = 80
            dl = DL([2] * R, R, 20)
            tdo = 1000
            x = rng.uniform([-1.0, -1.0, 0.0], [1.0, 1.0, 0.2], size=(R, 3))
            X, G, W, TS, AV, ST, GO, WO, TO = [], [], [], [], [], [], [], [], []
            for t in range(T):
                x = x + rng.normal(0, 0.02, size=(R, 3)) + 0.05 * (x.mean(axis=0) - x)
                x[:, 2] = np.abs(x[:, 2])
                goals = rng.uniform([-3, -3, 0.0], [3, 3, 0.3], size=(R, 3))
                weights = rng.choice([1.0, 2.0], size=R)
                avg = float(rng.choice([0.005, 0.02, 0.0299, 0.03, 0.031, 0.2]) * rng.uniform(0.95, 1.05))
                states = rng.choice([0, 1, 0, 1, 2], size=R)
                ts = t if rng.random() > 0.1 else int(rng.integers(0, 12))
                gl, wl = [g.copy() for g in goals], [float(w) for w in weights]
                X.append(x.copy()); G.append(goals.copy()); W.append(weights.copy()); TS.append(ts); AV.append(avg); ST.append(states)
                g_o, w_o, tdo = dl.deadlock_checking(x_robots=[xi.copy() for xi in x], goal_robots=gl, goal_weights=wl,
                                                     time_step=ts, time_deadlock_out=tdo, avg_sum=avg,
                                                     state_machine_robots=list(states))
                GO.append(np.array([np.asarray(g, dtype=np.float64) for g in g_o]))
                WO.append(np.array([float(w) for w in w_o]))
                TO.append(tdo)
            pre = f"c{case}_"
            out[pre + "R"] = np.array(R)
            for k, v in (("x", X), ("goals", G), ("weights", W), ("time_step", TS), ("avg", AV), ("states", ST),
                         ("goals_out", GO), ("weights_out", WO), ("tdo_out", TO)):
                out[pre + k] = np.array(v)
            case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "deadlock_point_golden.npz"), **out)
    print("wrote", case, "point-mass cases")


if __name__ == "__main__":
    sys.exit(main())
