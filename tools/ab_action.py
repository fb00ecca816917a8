"""A/B timing of the executed-action kernel (developer tool).  usage: python tools/ab_action.py [lib.so ...]
Shapes: C2 (2 Pandas, n = 4 spheres per link, S = 32) and C4 (3 Pandas, S = 64), 65 536 scenarios, FP32 and FP64."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import multi_robot_fabrics_b200 as m
    from multi_robot_fabrics_b200.api import Fabrics, to_soa
    B = int(os.environ.get("AB_B", 65536))
    out = {}
    for R in (2, 3):
        fab = Fabrics(R)
        rec = np.tile(m.scenarios.generate(4096, R, seed=1, weight_goal_1=20.0), (B // 4096, 1, 1))
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            d = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
            q, qd = d[0:7].contiguous(), d[7:14].contiguous()
            obst = fab.obstacles_dev(q, qd, n_per_link=4, vel_mode=0)
            act = torch.empty((7, R, B), dtype=dt, device="cuda:0")
            fn = lambda: fab.action_dev(d, obst, action=act)
            for _ in range(2): fn()
            torch.cuda.synchronize(); ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            S = obst.shape[0]
            ms = float(np.median(ts))
            out[f"R{R}_S{S}_{name}_ms"] = round(ms, 4)
            out[f"R{R}_S{S}_{name}_TF"] = round(B * R / (ms * 1e-3) * (5700 + 480 * S) / 1e12, 2)
            np.save(os.environ["AB_OUT"] + f".R{R}.{name}.npy", act[:, :, :4096].double().cpu().numpy())
        fab.close()
    print(json.dumps(out))
else:
    import numpy as np
    libs = [os.path.join(ROOT, "multi-robot-fabrics_b200", "libmrf_b200.so")] + sys.argv[1:]
    ref = {}
    for lib in libs:
        tag = os.path.basename(lib)[:-3]
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True, text=True,
                           env=dict(os.environ, AB_CHILD="1", MRF_B200_LIB=os.path.abspath(lib), AB_OUT=f"/tmp/aba_{tag}"))
        if p.returncode:
            print(tag, "FAILED", p.stderr[-400:]); continue
        res = json.loads(p.stdout.strip().split("\n")[-1])
        for f in sorted(glob.glob(f"/tmp/aba_{tag}.*.npy")):
            k = f.split(".", 1)[1]
            a = np.load(f)
            ref.setdefault(k, a)
            ok = np.isfinite(a) & np.isfinite(ref[k])
            res["maxdiff_" + k[:-4]] = float(np.abs(a - ref[k])[ok].max())
        print(tag, json.dumps(res), flush=True)
