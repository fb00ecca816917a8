"""Cooperative-kernel latency breakdown by configuration (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
def lat(fab, d, N, R):
    a = torch.empty((R, 1), dtype=d.dtype, device="cuda:0")
    for _ in range(10): fab.rollout_dev(d, N, avg_vel=a)
    torch.cuda.synchronize(); ts = []
    for _ in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fab.rollout_dev(d, N, avg_vel=a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))
for R in (2, 3):
    base = m.scenarios.generate(4, R, seed=0)[:1]
    d = torch.from_numpy(to_soa(base)).to("cuda:0", dtype=torch.float32)
    for name, kw in (("full RF-CV", dict(estimate_goal=1)), ("no goal estimate", dict()), ("no collision links", dict(has_collision_links=0))):
        fab = Fabrics(R, **kw)
        t20, t40 = lat(fab, d, 20, R), lat(fab, d, 40, R)
        print(f"R={R} {name}: H20 {t20:.1f} us, H40 {t40:.1f} us -> {(t40 - t20) / 20:.2f} us/step, fixed {t20 - (t40 - t20):.1f} us")
        fab.close()
