"""Kernel timeline of the pipelined RF-CV sweep (bench.py's _sweep) from CUPTI via torch.profiler: which kernel ran when,
on which stream.  Run on the GPU box:  python tools/timeline_probe.py [steps]  -> gpurun_out/timeline.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import multi_robot_fabrics_b200 as m  # noqa: E402
from multi_robot_fabrics_b200.api import Fabrics, to_soa  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R, B, H = 3, 65536, 20
dev = torch.device("cuda:0")
rec = m.scenarios.generate(B, R, seed=0).astype(np.float32)
fab = Fabrics(R, device=0, estimate_goal=1)
base = torch.from_numpy(to_soa(rec)).to(dev)
recs = [base] + [torch.roll(base, shifts=(k * B) // 6, dims=2).contiguous() for k in range(1, 6)]
works = [r.clone() for r in recs]
bench._sweep(fab, torch, None, dev, 1, recs, works, H, 4, 3, True)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = bench._sweep(fab, torch, None, dev, 1, recs, works, H, steps, 2, True)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline.txt"), "w") as f:
    f.write(f"total_ms {out['total_ms']:.4f} for {steps} steps\n")
    for e in ev:
        name = e.name[:60]
        f.write(f"{(e.time_range.start - t0):10.1f} us  +{e.time_range.elapsed_us():9.1f} us  {name}\n")
print(open(os.path.join(ROOT, "gpurun_out", "timeline.txt")).read()[:6000])
