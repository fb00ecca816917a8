"""Device-resident throughput of the rollout kernel for the BASELINE.json configurations (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
def timeit(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
B = 65536
for name, R, N in (("C3: 2 Pandas RF-CV H20", 2, 20), ("C4/C5: 3 Pandas RF-CV H50", 3, 50), ("metric shape: 3 Pandas RF-CV H20", 3, 20)):
    rec = np.tile(m.scenarios.generate(4096, R, seed=0), (B // 4096, 1, 1))
    fab = Fabrics(R, estimate_goal=1)
    for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
        d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
        a = torch.empty((R, B), dtype=dt, device="cuda:0")
        ms = timeit(lambda: fab.rollout_dev(d, N, avg_vel=a))
        S = 8 * (R - 1)
        rs = B * R * N / (ms * 1e-3)
        print(json.dumps(dict(config=name, dtype=nm, scenarios=B, ms=round(ms, 3), robot_steps_per_s=rs,
                              tflops_alg=rs * (5700 + 480 * S) / 1e12)))
    fab.close()
# decoupled (Cartesian-example) rollouts: one robot against the other robots' constant-velocity spheres, n = 4 per link
for name, R, N in (("FabricsRollouts, 2 Pandas, S = 32, H20", 2, 20), ("FabricsRollouts, 3 Pandas, S = 64, H50", 3, 50)):
    rec = np.tile(m.scenarios.generate(4096, R, seed=3, weight_goal_1=20.0), (B // 4096, 1, 1))
    fab = Fabrics(R)
    for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
        d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
        obst = fab.obstacles_dev(d[0:7].contiguous(), d[7:14].contiguous(), n_per_link=4, vel_mode=1)
        o0, r0 = obst[:, :, 0, :].contiguous(), d[:, 0, :].contiguous()
        a = torch.empty((B,), dtype=dt, device="cuda:0")
        ms = timeit(lambda: fab.rollout_cart_dev(0, r0, o0, N, avg_vel=a), n=5, warm=2)
        S = o0.shape[0]
        rs = B * N / (ms * 1e-3)
        print(json.dumps(dict(config=name, dtype=nm, scenarios=B, ms=round(ms, 3), robot_steps_per_s=rs,
                              tflops_alg=rs * (5700 + 480 * S) / 1e12)))
    fab.close()
