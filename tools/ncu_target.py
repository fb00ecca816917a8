"""Target for `ncu -k regex:rollout_kernel`: a few launches of the FP32 throughput rollout kernel on the bench shape
(65 536 x 3 Pandas x H20, RF-CV, stiffness output), nothing else of interest on the GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_robot_fabrics_b200 as m  # noqa: E402
from multi_robot_fabrics_b200.api import Fabrics, to_soa  # noqa: E402

B = int(os.environ.get("NCU_B", 65536))
dt = torch.float64 if os.environ.get("NCU_F64") else torch.float32
base = m.scenarios.generate(8192, 3, seed=0)
rec = np.tile(base, (B // 8192, 1, 1))
fab = Fabrics(3, estimate_goal=1)
d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
t = lambda *s: torch.empty(s, dtype=dt, device="cuda:0")
avg, xee, gest, risk = t(3, B), t(3, 3, B), t(3, B), t(3, B)
for _ in range(int(os.environ.get("NCU_LAUNCHES", 3))):
    fab.rollout_dev(d, 20, avg_vel=avg, x_ee=xee, goal_est=gest, risk=risk)
torch.cuda.synchronize()
