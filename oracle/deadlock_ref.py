"""Oracle for the deadlock heuristic -- TEST INFRASTRUCTURE ONLY.

A plain-Python restatement of deadlockprevention.deadlock_checking
(/root/reference/multi_robot_fabrics/others_planner/deadlock_prevention.py:6-34 constructor constants,
:50-118 the check).  PINNED: tests/golden/deadlock_golden.npz holds input/output sequences produced by the
reference's own class imported from /root/reference (generator: tests/golden/make_deadlock_golden.py), and
tests/test_oracle.py replays them through this restatement bit-for-bit.
"""
from __future__ import annotations

import itertools

import numpy as np


class DeadlockOracle:
    def __init__(self, n_robots: int, dist_endeff: float = 0.35, point: bool = False):
        self.n = n_robots
        self.dist_endeff = dist_endeff   # :64 (a literal in the reference; a knob here so tests can provoke deadlocks)
        if point:   # deadlock_prevention.py:12-19 (dof[0] == 2, point masses)
            self.avg_vel_constant, self.dist_constant = 0.03, 1
            self.w_follower, self.w_leader = 10, 1
            self.time_wait, self.goal_scale = 50, 100
        else:       # :20-27 (manipulator branch)
            self.avg_vel_constant, self.dist_constant = 0.16, 0.0
            self.w_follower, self.w_leader = 2, 3
            self.time_wait, self.goal_scale = 300, 2
        self.goal_robot0 = np.zeros(3)
        self.combos = list(itertools.combinations(range(n_robots), 2))   # :29-30
        self.i_leader, self.i_follower, self.dead = 0, 1, [0, 1]          # :10-11,33

    def step(self, x, goals, weights, time_step, tdo, avg_sum, states):
        """Returns (goals, weights, tdo, flag); goals / weights are updated copies."""
        x = [np.asarray(v, dtype=np.float64) for v in x]
        goals = [np.asarray(g, dtype=np.float64).copy() for g in goals]
        weights = list(weights)
        dist_goal = [np.linalg.norm(x[i] - goals[i]) for i in range(self.n)]          # :58
        flag, best = False, 100.0
        for (a, b) in self.combos:                                                    # :60
            d_ee = np.linalg.norm(x[a] - x[b])                                        # :63
            ok_state = states[a] in (0, 1) and states[b] in (0, 1)                    # :62
            if (avg_sum < self.avg_vel_constant and dist_goal[a] + dist_goal[b] > self.dist_constant
                    and time_step > 10 and ok_state and d_ee < self.dist_endeff):                 # :66
                flag = True                                                           # :73
                if d_ee < best:          # :74-80 rescans recorded distances with a strict '<': first minimum wins
                    best, self.dead = d_ee, [a, b]
        if flag and time_step > 10:                                                   # :83
            d0, d1 = self.dead
            if dist_goal[d0] > dist_goal[d1]:                                         # :85-90
                self.i_leader, self.i_follower = d1, d0
            else:
                self.i_leader, self.i_follower = d0, d1
            diff = x[self.i_leader] - x[self.i_follower]                              # :93
            diff_goal = diff * self.goal_scale
            nrm = np.linalg.norm(diff_goal)
            if nrm > 0.05:                                                            # :95-98
                g0 = x[self.i_follower] - 0.3 / nrm * diff_goal
            else:
                g0 = x[self.i_follower] - diff * self.goal_scale
            if g0[2] < 0:                                                             # :99-100
                g0[2] = 0.1
            self.goal_robot0 = g0
            weights[self.i_leader], weights[self.i_follower] = self.w_leader, self.w_follower   # :102-103
            goals[self.i_follower] = g0.copy()                                        # :104
            tdo = 0                                                                   # :106
        elif states[self.dead[0]] == 2 or states[self.dead[1]] == 2:                  # :108-109
            tdo = 400
        elif tdo < self.time_wait:                                                    # :111-115
            weights[self.i_leader], weights[self.i_follower] = self.w_leader, self.w_follower
            goals[self.i_follower] = self.goal_robot0.copy()
            tdo = tdo + 1
        return goals, weights, tdo, flag
