"""Generates tests/golden/fsm_golden.npz by driving the REFERENCE's own StateMachine class (imported from
/root/reference; numpy only) with seeded synthetic observations.  Run in the build container only."""
import importlib.util
import io
import os
import contextlib

import numpy as np

REF = "/root/reference/multi_robot_fabrics/others_planner/state_machine.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_state_machine", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.StateMachine


def main():
    SM = load_reference()
    rng = np.random.default_rng(11)
    out, case = {}, 0
    for rep in range(8):
        start = rng.uniform([0.2, -0.6, 1.1], [0.8, 0.6, 1.2])
        nr_blocks = int(rng.integers(1, 4))
        holder = {"x": start.copy()}
        sm = SM(start_goal=start.copy(), nr_robots=2, nr_blocks=nr_blocks, fk_fun_ee=lambda q: holder["x"],
                robot_types=["panda", "panda"])
        X, QG, GB, ST, GO, WE, GA = [], [], [], [], [], [], []
        x = start + rng.normal(0, 0.1, 3)
        qg = np.array([0.04, 0.04])
        block = rng.uniform([0.3, -0.2, 0.75], [0.7, 0.2, 0.8])
        for t in range(700):
            # a crude follower: the hand moves towards the machine's current goal, the gripper follows its action
            goal = np.asarray(sm.get_goal_robot(), dtype=np.float64)
            x = x + 0.08 * (goal - x) + rng.normal(0, 0.0005, 3)
            holder["x"] = x.copy()
            gb = block.copy()
            if sm.state_machine_panda in (12, 4, 5) or (sm.state_machine_panda == 3 and sm.time_gripping_panda > 20):
                gb = x - np.array([0.0, 0.0, 0.0])           # the block rides with the hand
            if rng.random() < 0.002:
                gb[2] = 0.5                                    # "dropped" block
            with contextlib.redirect_stdout(io.StringIO()):
                st = sm.get_state_machine_panda(q_robot=np.zeros(7), q_robot_gripper=qg.copy(), goal_block=gb.copy(),
                                                robot_type="panda")
            ga = sm.get_gripper_action_panda(qg.copy())
            X.append(x.copy()); QG.append(qg.copy()); GB.append(gb.copy()); ST.append(st)
            GO.append(np.asarray(sm.get_goal_robot(), dtype=np.float64).copy()); WE.append(float(sm.get_weight_goal0()))
            GA.append(ga.copy())
            qg = np.clip(qg + 0.01 * ga, 0.0, 0.04)
            if st == 0 and rng.random() < 0.05:
                block = rng.uniform([0.3, -0.2, 0.75], [0.7, 0.2, 0.8])
        pre = f"c{case}_"
        out[pre + "start"], out[pre + "nr_blocks"] = start, np.array(nr_blocks)
        for k, v in (("x", X), ("qg", QG), ("gb", GB), ("state", ST), ("goal", GO), ("weight", WE), ("grip", GA)):
            out[pre + k] = np.array(v)
        print("case", case, "states visited", sorted(set(ST)), "picked", sm.get_nr_blocks_picked())
        case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "fsm_golden.npz"), **out)


if __name__ == "__main__":
    main()
