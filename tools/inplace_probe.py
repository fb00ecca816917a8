"""Three blocking mrf_rollout_host_f32 calls on page-locked buffers (bench workload) -- target for an ncu capture of the
in-place (AOS) instance of rollout_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics
B, R, N = 65536, 3, 20
pin = lambda shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
rec = pin((B, R, 44)); rec[:] = np.tile(m.scenarios.generate(4096, R, seed=0).astype(np.float32), (B // 4096, 1, 1))
out = {"avg_vel": pin((B, R)), "x_ee": pin((B, R, 3)), "goal_est": pin((B, 3))}
fab = Fabrics(R, estimate_goal=1)
for _ in range(3):
    fab.rollout_host(rec, N, dtype="f32", out=out)
print("kernel ms", fab.handle.last_kernel_ms, "finite", float(np.isfinite(out["avg_vel"]).mean()))
