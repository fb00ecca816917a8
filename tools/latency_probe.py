"""Single-scenario rollout latency probe (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
R = int(os.environ.get("LP_R", 3)); N = int(os.environ.get("LP_N", 20)); B = int(os.environ.get("LP_B", 1))
base = m.scenarios.generate(max(B, 4), R, seed=0)[:B]
fab = Fabrics(R, estimate_goal=1)
for dt in (torch.float32, torch.float64):
    d = torch.from_numpy(to_soa(base)).to("cuda:0", dtype=dt)
    a = torch.empty((R, B), dtype=dt, device="cuda:0")
    for _ in range(10): fab.rollout_dev(d, N, avg_vel=a)
    torch.cuda.synchronize()
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fab.rollout_dev(d, N, avg_vel=a); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t0 = time.perf_counter()
    for _ in range(200): fab.rollout_dev(d, N, avg_vel=a)
    torch.cuda.synchronize()
    print(f"{dt} R={R} N={N} B={B}: events median {np.median(ts):.1f} us min {min(ts):.1f} us; back-to-back {(time.perf_counter()-t0)/200*1e6:.1f} us")
