"""Quick device-side timing of the rollout / action kernels (developer tool; bench.py is the contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

def main():
    B = int(os.environ.get("QB_B", 65536))
    base = m.scenarios.generate(4096, 3, seed=0)
    rec = np.tile(base, (B // 4096, 1, 1))
    fab = Fabrics(3, estimate_goal=1)
    print("fma peak f32 TF/s", fab.handle.fma_peak_tflops(False), "f64", fab.handle.fma_peak_tflops(True))
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
        for N in (20, 50):
            avg = torch.empty((3, B), dtype=dt, device="cuda:0")
            best, med = timeit(lambda: fab.rollout_dev(d, N, avg_vel=avg))
            rs = B * 3 * N / (best * 1e-3)
            print(json.dumps(dict(kernel="rollout", dtype=name, B=B, N=N, ms_best=best, ms_med=med, robot_steps_per_s=rs,
                                  tflops_alg=rs * 13.4e3 / 1e12)))
    # executed action with staged obstacles (configs C2 / C4 shapes: n = 4 spheres per link)
    for Rr, Bb in ((2, 65536), (3, 65536)):
        fabr = Fabrics(Rr)
        recr = np.tile(m.scenarios.generate(4096, Rr, seed=1, weight_goal_1=20.0), (Bb // 4096, 1, 1))
        dr = torch.from_numpy(to_soa(recr)).to("cuda:0", dtype=torch.float32)
        q, qd = dr[0:7].contiguous(), dr[7:14].contiguous()
        obst = fabr.obstacles_dev(q, qd, n_per_link=4, vel_mode=0)
        act = torch.empty((7, Rr, Bb), dtype=torch.float32, device="cuda:0")
        t_ob, _ = timeit(lambda: fabr.obstacles_dev(q, qd, n_per_link=4, vel_mode=0, obst=obst))
        t_ac, _ = timeit(lambda: fabr.action_dev(dr, obst, action=act))
        S = obst.shape[0]
        print(json.dumps(dict(kernel="obstacles+action", robots=Rr, B=Bb, S=S, ms_obstacles=t_ob, ms_action=t_ac,
                              actions_per_s=Bb * Rr / (t_ac * 1e-3), tflops_alg=Bb * Rr / (t_ac * 1e-3) * (5700 + 480 * S) / 1e12,
                              obst_GBps=obst.numel() * 4 / (t_ob * 1e-3) / 1e9)))
        fabr.close()
    # single-scenario latency
    d1 = torch.from_numpy(to_soa(base[:1])).to("cuda:0", dtype=torch.float32)
    avg1 = torch.empty((3, 1), dtype=torch.float32, device="cuda:0")
    best, med = timeit(lambda: fab.rollout_dev(d1, 20, avg_vel=avg1), n=20, warm=5)
    print(json.dumps(dict(kernel="rollout_single", N=20, us_best=best * 1e3, us_med=med * 1e3)))
    t0 = time.perf_counter()
    for _ in range(100):
        fab.rollout_dev(d1, 20, avg_vel=avg1)
    torch.cuda.synchronize()
    print("single rollout wall us (100 back-to-back)", (time.perf_counter() - t0) / 100 * 1e6)

if __name__ == "__main__":
    main()
