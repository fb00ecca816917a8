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
# round 2: stiffness output + RF-CV post step (guard select, strided FP64 re-roll through the list, deadlock with overrides),
# on two streams with two scratch slots; static spheres; collision-link subsets (generic + cooperative kernels)
d32 = d.float()
d32[7:14] *= 0.25
Bq = d32.shape[-1]
t = lambda *s_, dt=torch.float32: torch.zeros(s_, dtype=dt, device="cuda:0")
fab2 = Fabrics(R, estimate_goal=1, dl_dist_endeff=1.5)
fab2.set_guard(bands=[0.05, 0.05, 0.05, 0.5, 0.5, 0.5], cap=64)           # list (almost) everything: overflow path too
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
outs = []
for k, st in enumerate(streams):
    with torch.cuda.stream(st):
        avg, xee, gest, risk, res = t(R, Bq), t(R, 3, Bq), t(3, Bq), t(R, Bq), t(R + 1, Bq)
        work = d32.clone()
        fab2.rollout_dev(d32, N, avg_vel=avg, x_ee=xee, goal_est=gest, risk=risk)
        fl = fab2.rfcv_post_dev(d32, N, xee, work, gest, avg, t(R, Bq, dt=torch.int32), torch.full((Bq,), 100, dtype=torch.int32, device="cuda:0"),
                                torch.full((Bq,), 1000, dtype=torch.int32, device="cuda:0"),
                                torch.tensor([0, 1, 0, 1], dtype=torch.int32, device="cuda:0").repeat_interleave(Bq).contiguous(),
                                t(3, Bq), risk=risk, result=res, slot=k)
        outs.append((fl, res))
torch.cuda.synchronize()
print("guard stats", fab2.guard_stats())
stat = t(3, 4, R, Bq)
stat[:, 0:3] = torch.rand((3, 3, R, Bq), device="cuda:0") + 1.5
stat[:, 3] = 0.08
fab2.rollout_static_dev(d32, stat, N)
fab3 = Fabrics(2, collision_links=[[5], [3, 6, 8]], r_robots=[[0.08, 0.06, 0.08, 0.08, 0.07, 0.09, 0.08, 0.08]] * 2)
rec2 = m.scenarios.generate(40, 2, seed=2)
for coop in (0, 1 << 20):
    fab3.handle.set_coop_max_batch(coop)
    fab3.rollout_host(rec2, N, dtype="f32", trajectories=True)
fab3.action_host(rec2, np.random.default_rng(0).uniform(1.5, 2.0, size=(40, 2, 3, 10)), dtype="f32")
torch.cuda.synchronize()
print("sanitize probe done", float(act.abs().max()))
