"""Closed-loop batched episodes: the control loop of examples/example_pandas_Jointspace.py:280-458 for thousands of
independent scenarios at once, entirely on the GPU (no pybullet).

Per control step, for every scenario (file:line in examples/example_pandas_Jointspace.py):
  1. end-effector FK and the RF-CV goal estimate of robot 1                :324-348   (rollout kernel pre-pass)
  2. coupled Rollout-Fabrics horizon -> mean-square velocity per robot     :352-376   (rollout kernel, weight_goal_1 = 10)
  3. deadlock_checking -> follower goal / weights / timer                   :379-385   (deadlock kernel)
  4. other robots' collision spheres (n_obst_per_link, link-origin velocity):400-412   (obstacles kernel)
  5. executed action per robot (weight_goal_1 = 20), velocity clip          :417-453   (action kernel)
  6. kinematic environment step q += dt * clip(action)  (urdfenvs 'vel' mode stand-in for pybullet, :454)
The pick-and-place state machine is replaced by a reach task: every robot keeps its goal, state-machine code 0
("move to goal"), and an episode succeeds when every hand is within `epsilon` (0.05, the goal struct's epsilon, :35)
of its own goal.  Torch is used for tensor hand-off and slice bookkeeping only; one control step is captured in a CUDA
graph and replayed.
"""
from __future__ import annotations

import numpy as np

from ._lib import G0, Q, QD, W0, W1
from .api import Fabrics, to_soa

VEL_LIMITS = [2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61]          # example_pandas_Jointspace.py:221


class BatchedEpisodes:
    def __init__(self, rec: np.ndarray, n_horizon: int = 10, rollout_fabrics: bool = True, resolve_deadlocks: bool = True,
                 estimate_goal: bool = False, static_or_dyn: int = 1, n_obst_per_link: int = 1, device: int = 0,
                 dtype: str = "f32", epsilon: float = 0.05, use_graph: bool = True):
        import torch
        self.torch = torch
        B, R, _ = rec.shape
        self.B, self.R, self.N = B, R, int(n_horizon)
        self.rollout_fabrics, self.resolve_deadlocks = rollout_fabrics, resolve_deadlocks
        self.n_per_link, self.epsilon, self.use_graph = int(n_obst_per_link), float(epsilon), use_graph
        self.dev = torch.device(f"cuda:{device}")
        self.tdt = torch.float32 if dtype == "f32" else torch.float64
        self.fab = Fabrics(R, device=device, estimate_goal=1 if estimate_goal else 0, static_or_dyn=static_or_dyn)
        self.fab.handle.set_coop_max_batch(0 if B > 256 else 1 << 20)
        t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, dtype=dt or self.tdt)
        self.rec = t(to_soa(rec))                                    # (44,R,B): q, qdot rows are the live state
        self.goal0 = self.rec[G0:G0 + 3].permute(1, 0, 2).contiguous().clone()     # (R,3,B) task goals
        self.w0 = self.rec[W0].clone()                                              # (R,B)
        self.goals = self.goal0.clone()                                             # working copies (deadlock mutates)
        self.weights = self.w0.clone()
        z = lambda *s, dt=None: torch.zeros(s, dtype=dt or self.tdt, device=self.dev)
        self.avg, self.xee, self.gest = z(R, B), z(R, 3, B), z(3, B)
        self.obst = z(8 * self.n_per_link * (R - 1), 10, R, B)
        self.act = z(7, R, B)
        self.sm = z(R, B, dt=torch.int32)                                            # state-machine code 0
        self.tstep = z(B, dt=torch.int32)
        self.tdo = torch.full((B,), 1000, dtype=torch.int32, device=self.dev)       # :275
        self.st_int = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=self.dev).repeat_interleave(B).contiguous()
        self.st_goal = z(3, B)
        self.flag = z(B, dt=torch.int32)
        self.vlim = torch.tensor(VEL_LIMITS, dtype=self.tdt, device=self.dev).view(7, 1, 1)
        self.dt = float(self.fab.cfg.dt)
        # metrics
        self.done_at = torch.full((B,), -1, dtype=torch.int32, device=self.dev)
        self.deadlock_steps = z(B, dt=torch.int32)
        self.min_clear = torch.full((B,), 100.0, dtype=self.tdt, device=self.dev)   # :231 min_clearance = 100
        self.sx = z(8 * self.n_per_link, 3, R, B)
        self.kx, self.kv, self.ka = z(8, 3, R, B), z(8, 3, R, B), z(8, 3, R, B)
        self.graph = None
        self.steps_done = 0

    # one control step; every operation is a kernel launch on the current stream (graph-capturable)
    def _step(self):
        torch, R = self.torch, self.R
        rec = self.rec
        rec[QD:QD + 7].copy_(torch.minimum(torch.maximum(rec[QD:QD + 7], -self.vlim), self.vlim))     # :288
        self.goals.copy_(self.goal0)
        self.weights.copy_(self.w0)
        if self.rollout_fabrics:
            rec[G0:G0 + 3].copy_(self.goals.permute(1, 0, 2))
            rec[W0].copy_(self.weights)
            rec[W1].fill_(10.0)                                                                       # :43,364
            self.fab.rollout_dev(rec, self.N, avg_vel=self.avg, x_ee=self.xee, goal_est=self.gest)
            if self.fab.cfg.estimate_goal and R > 1:
                self.goals[1].copy_(self.gest)                                                        # :346-348
            if self.resolve_deadlocks:
                self.fab.deadlock_dev(self.xee, self.goals, self.weights, self.sm, self.tstep, self.tdo, self.st_int,
                                      self.st_goal, avg_vel=self.avg, flag=self.flag)
                self.deadlock_steps.add_(self.flag)
        else:
            self.fab.kinematics_dev(rec[Q:Q + 7], rec[QD:QD + 7], x=self.kx, v=self.kv, a=self.ka)    # MRDF: hands' FK only
            self.xee.copy_(self.kx[7].permute(1, 0, 2))
        # executed action with the (possibly overridden) goals / weights, weight_goal_1 = 20          :417-445
        rec[G0:G0 + 3].copy_(self.goals.permute(1, 0, 2))
        rec[W0].copy_(self.weights)
        rec[W1].fill_(20.0)
        q, qd = rec[Q:Q + 7], rec[QD:QD + 7]
        self.fab.obstacles_dev(q, qd, n_per_link=self.n_per_link, vel_mode=0, obst=self.obst, spheres_x=self.sx)
        self.fab.action_dev(rec, self.obst, action=self.act)
        a = torch.minimum(torch.maximum(self.act, -self.vlim), self.vlim)                             # :453
        a = torch.where(torch.isfinite(a), a, torch.zeros_like(a))
        qd.copy_(a)
        q.add_(a * self.dt)                                                                           # kinematic env step
        # metrics: reach test on the task goals, minimum sphere clearance between robots (:461-470)
        dist = (self.xee - self.goal0).square().sum(dim=1).sqrt()                                     # (R,B)
        reached = (dist < self.epsilon).all(dim=0)
        first = reached & (self.done_at < 0)
        self.done_at.copy_(torch.where(first, self.tstep, self.done_at))
        for a_ in range(R):
            for b_ in range(a_ + 1, R):
                d = (self.sx[:, None, :, a_] - self.sx[None, :, :, b_]).square().sum(dim=2).sqrt()   # (S,S,B)
                self.min_clear.copy_(torch.minimum(self.min_clear, d.flatten(0, 1).min(dim=0).values - 0.16))
        self.tstep.add_(1)

    def run(self, n_steps: int):
        torch = self.torch
        if self.use_graph and self.graph is None:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(2):                      # warm-up outside the capture
                    self._step()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            self.steps_done += 2
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):          # capture records the launches, it does not execute them
                self._step()
            n_steps -= 2
        for _ in range(max(0, n_steps)):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step()
            self.steps_done += 1
        return self

    def results(self) -> dict:
        self.torch.cuda.synchronize(self.dev)
        done = self.done_at.cpu().numpy()
        return {"steps": self.steps_done, "success": done >= 0, "steps_to_success": done,
                "deadlock_steps": self.deadlock_steps.cpu().numpy(),
                "min_clearance": self.min_clear.double().cpu().numpy(),
                "q": self.rec[Q:Q + 7].permute(2, 1, 0).double().cpu().numpy(),
                "x_ee": self.xee.permute(2, 0, 1).double().cpu().numpy()}
