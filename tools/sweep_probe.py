"""What the pipelined RF-CV sweep (bench.py's _sweep) costs beyond its rollout kernels: the same sweep with the post step
removed, with other stream counts / priorities (developer tool).  Run on the GPU box: python tools/sweep_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import multi_robot_fabrics_b200 as m  # noqa: E402
from multi_robot_fabrics_b200.api import Fabrics, to_soa  # noqa: E402

R, B, H, K = 3, 65536, 20, 20
dev = torch.device("cuda:0")
rec = m.scenarios.generate(B, R, seed=0).astype(np.float32)
fab = Fabrics(R, device=0, estimate_goal=1)
base = torch.from_numpy(to_soa(rec)).to(dev)
recs = [base] + [torch.roll(base, shifts=(k * B) // 6, dims=2).contiguous() for k in range(1, 6)]
works = [r.clone() for r in recs]
real_post = fab.rfcv_post_dev


def run(tag, post=True, risk=True, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    fab.rfcv_post_dev = real_post if post else (lambda *a, **k: None)
    bench._sweep(fab, torch, None, dev, 1, recs, works, H, 5, 3, risk)
    ts = [bench._sweep(fab, torch, None, dev, 1, recs, works, H, K, 5, risk)["total_ms"] / K for _ in range(3)]
    print(f"{tag:40s} ms/step {min(ts):.4f} (runs {' '.join(f'{t:.4f}' for t in ts)})", flush=True)
    for k in env:
        os.environ.pop(k)


tag = os.environ.get("TAG", "")
run(tag + "post step, normal priority, 8 output sets (bench default)")
run(tag + "no post step", post=False)
run(tag + "no post step, 1 rollout stream", post=False, MRF_BENCH_NROLL=1)
run(tag + "post step, high priority, 4 output sets", MRF_BENCH_PRIO=-1, MRF_BENCH_NBUF=4)
run(tag + "post step, normal priority, 4 output sets", MRF_BENCH_NBUF=4)
run(tag + "post = heuristic kernel only (no guard)", risk=False)
tiny = torch.zeros(1024, device=dev)
for n in (1, 6):
    def only_tiny(*a, **k):
        for _ in range(n):
            tiny.add_(1.0)          # one 1-CTA kernel on the post stream
    fab.rfcv_post_dev = only_tiny
    bench._sweep(fab, torch, None, dev, 1, recs, works, H, 5, 3, True)
    ts = [bench._sweep(fab, torch, None, dev, 1, recs, works, H, K, 5, True)["total_ms"] / K for _ in range(3)]
    print(f"{tag}post = {n} empty kernel(s)                ms/step {min(ts):.4f}", flush=True)
