"""End-to-end (host buffers) timing of mrf_rollout_host for the bench workload: staged chunk pipeline vs the kernel
reading page-locked records in place, over admission-window sizes (developer tool)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics

def main():
    B, R, N = int(os.environ.get("QB_B", 65536)), 3, 20
    base = m.scenarios.generate(4096, R, seed=0).astype(np.float32)
    pin = lambda shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
    rec = pin((B, R, 44)); rec[:] = np.tile(base, (B // 4096, 1, 1))
    out = {"avg_vel": pin((B, R)), "x_ee": pin((B, R, 3)), "goal_est": pin((B, 3))}
    for zc, win in ((0, 0), (1, 8), (1, 16), (1, 32), (1, 64), (1, 128), (1, 100000)):
        os.environ["MRF_ZERO_COPY"] = str(zc)
        os.environ["MRF_ZC_WINDOW"] = str(max(win, 1))
        fab = Fabrics(R, estimate_goal=1)
        for _ in range(3):
            fab.rollout_host(rec, N, dtype="f32", out=out)
        ts, ks = [], []
        for _ in range(15):
            t0 = time.perf_counter()
            fab.rollout_host(rec, N, dtype="f32", out=out)
            ts.append(time.perf_counter() - t0)
            ks.append(fab.handle.last_kernel_ms)
        t = float(np.median(ts))
        print(json.dumps(dict(zero_copy=zc, window=win, ms_wall_median=t * 1e3, ms_wall_min=min(ts) * 1e3,
                              ms_events=float(np.median(ks)), robot_steps_per_s=B * R * N / t)))
        fab.close()

if __name__ == "__main__":
    main()
