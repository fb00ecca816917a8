"""Rollout time vs batch size for both rollout kernels (3 Pandas, H = 20, FP32): where the cooperative low-latency
kernel hands over to the throughput kernel (developer tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
def timeit(fn, n=9, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
R, N = 3, 20
base = m.scenarios.generate(4096, R, seed=0)
fab = Fabrics(R, estimate_goal=1)
for B in (1, 32, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    rec = np.tile(base, ((B + 4095) // 4096, 1, 1))[:B]
    d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=torch.float32)
    a = torch.empty((R, B), dtype=torch.float32, device="cuda:0")
    out = {"B": B}
    for name, thr in (("cooperative", 1 << 30), ("throughput", 0)):
        fab.handle.set_coop_max_batch(thr)
        ms = timeit(lambda: fab.rollout_dev(d, N, avg_vel=a))
        out[name + "_us"] = round(ms * 1e3, 1)
    out["robot_steps_per_s_best"] = B * R * N / (min(out["cooperative_us"], out["throughput_us"]) * 1e-6)
    print(json.dumps(out))
