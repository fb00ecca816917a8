"""Are the host-record (AOS) and device-SoA instantiations of the FP32 rollout kernel bitwise identical?  (developer tool)
usage: python tools/ab_layouts.py lib.so [...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import multi_robot_fabrics_b200 as m
    from multi_robot_fabrics_b200.api import Fabrics, to_soa
    B, N = 8192, 20
    rec = m.scenarios.generate(B, 3, seed=5).astype(np.float32)
    fab = Fabrics(3, estimate_goal=1)
    fab.handle.set_coop_max_batch(0)
    pin = torch.from_numpy(rec).pin_memory().numpy()
    h = fab.rollout_host(pin, N, dtype="f32")
    d = torch.from_numpy(to_soa(rec)).to("cuda:0")
    avg = torch.empty((3, B), dtype=torch.float32, device="cuda:0")
    fab.rollout_dev(d, N, avg_vel=avg)
    a = avg.T.contiguous().cpu().numpy()
    ok = np.isfinite(a) & np.isfinite(h["avg_vel"])
    diff = (a.view(np.uint32) != h["avg_vel"].view(np.uint32)) & ok
    print("differing", int(diff.sum()), "of", a.size, "max abs", float(np.abs(a - h["avg_vel"])[ok].max()))
else:
    for lib in sys.argv[1:]:
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, AB_CHILD="1", MRF_B200_LIB=os.path.abspath(lib)),
                           capture_output=True, text=True)
        print(os.path.basename(lib), p.stdout.strip().split("\n")[-1] if p.returncode == 0 else p.stderr[-300:])
