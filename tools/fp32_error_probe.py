"""FP32-path error against the float64 oracle over the horizon (developer tool; numbers quoted in DESIGN.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics
from helpers import oracle_rollout
for R, N in ((2, 20), (3, 20), (3, 50)):
    B = 4096
    rec = m.scenarios.generate(B, R, seed=99)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N)
    fab = Fabrics(R)
    for kernel, cm in (("throughput", 0), ("cooperative", 1 << 20)):
        fab.handle.set_coop_max_batch(cm)
        out = fab.rollout_host(rec, N, dtype="f32", trajectories=True)
        fin = np.isfinite(out["qdN"]).all(axis=(1, 2, 3))
        sel = ok & fin
        eqd = np.abs(out["qdN"] - qdN)[sel].max(axis=(1, 2, 3))
        eq = np.abs(out["qN"] - qN)[sel].max(axis=(1, 2, 3))
        ea = np.abs(out["avg_vel"] - avg)[sel].max(axis=1)
        print(f"R={R} N={N} {kernel}: scenarios ok {ok.sum()}/{B}, f32 finite among ok {(fin & ok).sum()}; "
              f"|dqdot| max {eqd.max():.2e} p99 {np.quantile(eqd, 0.99):.2e} median {np.median(eqd):.2e}; "
              f"|dq| max {eq.max():.2e}; |davg_vel| max {ea.max():.2e}")
    fab.close()
