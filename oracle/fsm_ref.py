"""Oracle for the pick-and-place state machine -- TEST INFRASTRUCTURE ONLY.

Plain-Python restatement of StateMachine.get_state_machine_panda / get_gripper_action_panda
(/root/reference/multi_robot_fabrics/others_planner/state_machine.py:4-37 constants, :70-84 gripper, :133-214 FSM).
PINNED: tests/golden/fsm_golden.npz holds sequences produced by the reference's own class imported from
/root/reference (generator tests/golden/make_fsm_golden.py); tests replay them through this restatement exactly.
"""
from __future__ import annotations

import numpy as np


class FsmOracle:
    def __init__(self, start_goal, nr_blocks):
        self.state, self.n_ok, self.n_fail, self.nr_blocks = 1, 0, 0, nr_blocks      # :8-11
        self.weight, self.closed, self.t_grip, self.stop = 2, False, 0, 0            # :12,30,37
        self.start = np.asarray(start_goal, dtype=np.float64)
        self.goal = self.start.copy()
        self.above = np.zeros(3)

    def step(self, x_ee, q_grip, goal_block):
        x = np.asarray(x_ee, dtype=np.float64)
        gb = np.asarray(goal_block, dtype=np.float64)
        pre = gb + np.array([0.0, 0.0, 0.1])                                          # :134-135
        d_start = np.linalg.norm(x - self.start)
        d_pre = np.linalg.norm(x[:2] - pre[:2])
        d_block = np.linalg.norm(x - gb)
        d_open = np.linalg.norm(np.asarray(q_grip, dtype=np.float64) - np.array([0.04, 0.04]))
        if self.n_ok > self.nr_blocks - 1:                                            # :141-142
            self.state = 10
        elif gb[2] < 0.6:                                                             # :143-147
            self.n_ok += 1
            self.n_fail += 1
            self.state = 0
        s = self.state
        if s == 0:                                                                    # :150-155
            self.goal, self.closed = self.start.copy(), False
            if d_start < 0.05:
                self.state = 1
        elif s == 1:                                                                  # :158-162
            self.goal = pre
            if d_pre < 0.013:
                self.state = 2
        elif s == 2:                                                                  # :164-170
            self.goal = gb.copy()
            if d_block < 0.013:
                self.closed, self.weight, self.state = True, 0, 3
        elif s == 3:                                                                  # :172-181
            self.goal = gb.copy()
            self.above = gb + np.array([0.0, 0.0, 0.15])
            self.t_grip += 1
            if self.t_grip > 0.3 / 0.01:
                self.t_grip, self.goal, self.weight, self.state = 0, self.start.copy(), 2, 12
        elif s == 12:                                                                 # :183-186
            self.goal = self.above.copy()
            if np.linalg.norm(x[:2] - self.goal[:2]) < 0.04:
                self.state = 4
        elif s == 4:                                                                  # :188-194
            self.goal = self.start.copy()
            if d_start < 0.15:
                self.state, self.closed = 5, False
        elif s == 5:                                                                  # :196-200
            if d_open < 0.005:
                self.state, self.n_ok, self.goal = 0, self.n_ok + 1, self.start.copy()
        elif s == 10:                                                                 # :202-206
            self.stop = 1
        return self.state

    def gripper_action(self, q_grip):                                                 # :70-84
        q = np.asarray(q_grip, dtype=np.float64)
        a = np.zeros(2)
        if self.closed:
            a[:] = -0.05
        elif np.linalg.norm(q - np.array([0.04, 0.04])) > 0.005:
            for z in range(2):
                a[z] = -0.4 if q[z] > 0.04 else 0.4
        return a
