"""A/B timing of the FP32 (and optionally FP64) throughput rollout kernel over library builds (developer tool).
usage: python tools/ab_rollout.py [build_ab/*.so ...]   -- every library is timed in its own process on the bench shape
(65 536 x 3 Pandas x H20, RF-CV, risk output); avg_vel is compared with the first library's."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import multi_robot_fabrics_b200 as m
    from multi_robot_fabrics_b200.api import Fabrics, to_soa

    B = int(os.environ.get("AB_B", 65536))
    N = int(os.environ.get("AB_N", 20))
    base = m.scenarios.generate(8192, 3, seed=0)
    rec = np.tile(base, (B // 8192, 1, 1))
    fab = Fabrics(3, estimate_goal=1)
    out = {}
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        if name == "f64" and not os.environ.get("AB_F64"):
            continue
        d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
        t = lambda *s: torch.empty(s, dtype=dt, device="cuda:0")
        avg, xee, gest, risk = t(3, B), t(3, 3, B), t(3, B), t(3, B)
        fn = lambda: fab.rollout_dev(d, N, avg_vel=avg, x_ee=xee, goal_est=gest, risk=risk)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(9 if name == "f32" else 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name + "_ms_med"] = float(np.median(ts))
        out[name + "_ms_min"] = float(min(ts))
        a = avg[:, :8192].double().cpu().numpy()
        np.save(os.environ["AB_OUT"] + "." + name + ".npy", a)
    print(json.dumps(out))


def main():
    libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build_ab", "*.so")))
    libs = [os.path.join(ROOT, "multi-robot-fabrics_b200", "libmrf_b200.so")] + libs
    import numpy as np
    ref = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for lib in libs:
        tag = os.path.basename(lib)[:-3]
        env = dict(os.environ, MRF_B200_LIB=os.path.abspath(lib), AB_CHILD="1", AB_OUT=f"/tmp/ab_{tag}")
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(tag, "FAILED", p.stderr[-400:])
            continue
        res = json.loads(p.stdout.strip().split("\n")[-1])
        for name in ("f32", "f64"):
            f = f"/tmp/ab_{tag}.{name}.npy"
            if os.path.exists(f):
                a = np.load(f)
                if name not in ref:
                    ref[name] = a
                ok = np.isfinite(a) & np.isfinite(ref[name])
                if ok.any():
                    res[name + "_maxdiff"] = float(np.abs(a - ref[name])[ok].max())
                    res[name + "_p99diff"] = float(np.quantile(np.abs(a - ref[name])[ok], 0.99))
                res[name + "_finite"] = float(np.isfinite(a).mean())
        print(tag, json.dumps(res), flush=True)


if __name__ == "__main__":
    child() if os.environ.get("AB_CHILD") else main()
