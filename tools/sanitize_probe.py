"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck) covering every kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
R, N, B = 3, 4, 70
rec = m.scenarios.generate(B, R, seed=1)
fab = Fabrics(R, estimate_goal=1)
for coop in (0, 1 << 20):
    fab.handle.set_coop_max_batch(coop)
    for dt in ("f32", "f64"):
        out = fab.rollout_host(rec, N, dtype=dt, trajectories=True)
        assert np.isfinite(out["avg_vel"]).mean() > 0.9
d = torch.from_numpy(to_soa(rec)).to("cuda:0")
obst = fab.obstacles_dev(d[0:7].contiguous(), d[7:14].contiguous(), n_per_link=2, vel_mode=1)
act = fab.action_dev(d, obst)
fab.rollout_cart_dev(0, d[:, 0].contiguous(), obst[:, :, 0].contiguous(), N)
fab.kinematics_dev(d[0:7].contiguous(), d[7:14].contiguous())
fab.handle.set_coop_max_batch(512)
big = fab.rollout_host(np.tile(rec, (120, 1, 1))[:8200].astype(np.float32), 2, dtype="f32")   # pipelined host path
pinned = torch.from_numpy(np.tile(rec, (120, 1, 1))[:8229].astype(np.float32)).pin_memory().numpy()
inplace = fab.rollout_host(pinned, 2, dtype="f32")        # page-locked records read in place (ticketed tiles, ragged tail)
assert np.array_equal(inplace["avg_vel"][:8200].view(np.uint8), big["avg_vel"].view(np.uint8))
from multi_robot_fabrics_b200.episodes import BatchedEpisodes
blocks, start = m.scenarios.pick_and_place_layout(rec, n_blocks=1, seed=1)
for kw in (dict(rollout_fabrics=True, resolve_deadlocks=True, estimate_goal=True), dict(rollout_fabrics=False)):
    BatchedEpisodes(rec, n_horizon=3, dtype="f32", n_obst_per_link=2, use_graph=False, **kw).run(3).results()
    BatchedEpisodes(rec, n_horizon=3, dtype="f32", use_graph=False, blocks=blocks, start_goal=start, **kw).run(3).results()
torch.cuda.synchronize()
print("sanitize probe done", float(act.abs().max()))
