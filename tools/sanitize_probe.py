"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck) covering every kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
R, N, B = 3, 4, 70
rec = m.scenarios.generate(B, R, seed=1)
fab = Fabrics(R, estimate_goal=1)
for coop in (0, 1 << 20):
    fab.handle.set_coop_max_batch(coop)
    for dt in ("f32", "f64"):
        out = fab.rollout_host(rec, N, dtype=dt, trajectories=True)
        assert np.isfinite(out["avg_vel"]).mean() > 0.9
d = torch.from_numpy(to_soa(rec)).to("cuda:0")
obst = fab.obstacles_dev(d[0:7].contiguous(), d[7:14].contiguous(), n_per_link=2, vel_mode=1)
act = fab.action_dev(d, obst)
fab.rollout_cart_dev(0, d[:, 0].contiguous(), obst[:, :, 0].contiguous(), N)
fab.kinematics_dev(d[0:7].contiguous(), d[7:14].contiguous())
big = fab.rollout_host(np.tile(rec, (120, 1, 1))[:8200].astype(np.float32), 2, dtype="f32")   # pipelined host path
torch.cuda.synchronize()
print("sanitize probe done", float(act.abs().max()))
