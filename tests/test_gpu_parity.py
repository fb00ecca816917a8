"""GPU parity tests: the CUDA kernels, called through the C-ABI, against oracle O2 on the same seeded inputs.

Tolerances (north_star): FP64 within 1e-9 relative on actions / trajectories; FP32 within the tolerance stated in
each test; deadlock flags identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200.api import Fabrics, to_soa
from oracle import o2

from helpers import oracle_actions, oracle_rollout, random_obstacles, rel_ps

F64_RTOL = 1e-9


@pytest.fixture(scope="module")
def fabs(built):
    d = {}
    yield d
    for f in d.values():
        f.close()


def get_fab(fabs, R, **kw):
    key = (R, tuple(sorted(kw.items())))
    if key not in fabs:
        fabs[key] = Fabrics(R, device=0, **kw)
    return fabs[key]


@pytest.mark.parametrize("kernel", ["throughput", "cooperative"])
@pytest.mark.parametrize("R,N,B,est", [(2, 20, 70, 0), (2, 20, 33, 1), (3, 50, 96, 1), (3, 20, 1, 0), (3, 50, 40, 2)])
def test_rollout_f64_matches_oracle(fabs, R, N, B, est, kernel):
    rec = m.scenarios.generate(B, R, seed=100 + R + N)
    fab = get_fab(fabs, R, estimate_goal=est)
    fab.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    fab.handle.set_coop_max_batch(512)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N, estimate_goal=est)
    assert ok.sum() >= max(1, int(0.9 * B))
    # 1e-9 relative PER SCENARIO: every scenario's worst error against that scenario's own largest value
    rel = lambda got, ref, ax: (np.abs(got - ref).max(axis=ax) / np.abs(ref).max(axis=ax))[ok].max()
    assert rel(out["qdN"], qdN, (1, 2, 3)) < F64_RTOL
    assert rel(out["qN"], qN, (1, 2, 3)) < F64_RTOL
    assert rel(out["avg_vel"], avg, 1) < F64_RTOL
    assert np.abs(out["x_ee"] - xee).max() < 1e-12
    assert np.abs(out["goal_est"] - goal).max() < 1e-12


@pytest.mark.parametrize("kernel", ["throughput", "cooperative"])
@pytest.mark.parametrize("R,N", [(2, 20), (3, 50)])
def test_rollout_f32_within_tolerance(fabs, R, N, kernel):
    """FP32 path against the float64 oracle over the horizon.  Stated tolerance: per scenario the worst |qdot - oracle|
    over all robots / steps / joints is <= 2e-3 rad/s (and |q - oracle| <= 2e-4 rad, |avg_vel - oracle| <= 1e-3) for at
    least 99.5 % of random scenarios, with the 99th percentile below 1e-4 rad/s.  The remaining < 0.5 % are numerically
    stiff near-contact scenarios (leaf force ~ 1/x^8 under the reference's explicit dt = 0.01 integration) in which any
    rounding difference is amplified along the horizon; the FP64 path is the one with a per-scenario guarantee."""
    B = 1024
    rec = m.scenarios.generate(B, R, seed=7)
    fab = get_fab(fabs, R)
    fab.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    out = fab.rollout_host(rec, N, dtype="f32", trajectories=True)
    fab.handle.set_coop_max_batch(512)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N)
    with np.errstate(invalid="ignore"):
        eqd = np.nan_to_num(np.abs(out["qdN"] - qdN).max(axis=(1, 2, 3)), nan=np.inf)[ok]
        eq = np.nan_to_num(np.abs(out["qN"] - qN).max(axis=(1, 2, 3)), nan=np.inf)[ok]
        ea = np.nan_to_num(np.abs(out["avg_vel"] - avg).max(axis=1), nan=np.inf)[ok]
    assert ok.sum() > 0.95 * B
    assert np.mean(eqd <= 2e-3) >= 0.995 and np.mean(eq <= 2e-4) >= 0.995 and np.mean(ea <= 1e-3) >= 0.995
    assert np.quantile(eqd, 0.99) < 1e-4 and np.median(eqd) < 5e-6
    assert np.abs(out["x_ee"] - xee).max() < 1e-5


@pytest.mark.parametrize("kernel", ["throughput", "cooperative"])
def test_rollout_static_fabrics_and_nonuniform_radii(fabs, kernel):
    """STATIC_OR_DYN_FABRICS = 0 zeroes the other robots' v and a (forward_planner_Jointspace.py:215-217); unequal
    sphere radii disable the merged-point fast path (link1/link2 and link5/link6 become separate leaves)."""
    R, N, B = 2, 10, 32
    rec = m.scenarios.generate(B, R, seed=3)
    fab = get_fab(fabs, R, static_or_dyn=0)
    fab.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N, static_or_dyn=0)
    rel = lambda got, ref: (np.abs(got - ref).max(axis=(1, 2, 3)) / np.abs(ref).max(axis=(1, 2, 3)))
    assert rel(out["qdN"], qdN)[ok].max() < F64_RTOL
    rr = [[0.08, 0.06, 0.08, 0.08, 0.07, 0.09, 0.08, 0.08], [0.05, 0.05, 0.08, 0.1, 0.08, 0.08, 0.06, 0.08]]
    rec[:, :, o2.RB:o2.RB + 6] = [0.08, 0.07, 0.09, 0.06, 0.08, 0.1]
    fab2 = Fabrics(R, device=0, r_robots=rr)
    fab2.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    out = fab2.rollout_host(rec, N, dtype="f64", trajectories=True)
    fab2.close()
    ocfg = o2.default_config(R)
    for r in range(R):
        for l in range(8):
            ocfg.r_robots[r][l] = rr[r][l]
    qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec, N)
    ok = np.isfinite(qdN).all(axis=(1, 2, 3)) & (np.abs(qdN).max(axis=(1, 2, 3)) < 3)
    assert ok.sum() > B // 2
    assert rel(out["qdN"], qdN)[ok].max() < F64_RTOL


@pytest.mark.parametrize("n_rob,S", [(2, 32), (1, 0), (3, 64), (2, 5)])
def test_action_matches_oracle(fabs, n_rob, S):
    """Executed action (compute_action without the small-action clamp), S dynamic obstacle spheres."""
    B = 50
    rng = np.random.default_rng(S + n_rob)
    R = max(2, n_rob)
    rec = m.scenarios.generate(B, R, seed=11, weight_goal_1=20.0)[:, :n_rob]
    obst = random_obstacles(rng, B, n_rob, S, rec)
    fab = get_fab(fabs, R)
    act = fab.action_host(rec, obst if S else None, robot_first=0, dtype="f64")
    ref = oracle_actions(rec, obst)
    ok = np.isfinite(ref).all(axis=(1, 2))
    assert ok.sum() > 0.9 * B
    assert rel_ps(act, ref, ok) < F64_RTOL
    act32 = fab.action_host(rec, obst if S else None, robot_first=0, dtype="f32")
    assert np.abs(act32 - ref)[ok].max() < 2e-3


def test_action_nonuniform_body_radii(fabs):
    """Executed action with a different radius on every ego link (link5 != link6: two leaves at one point) -- the
    obstacle-major accumulation of both precisions must treat the shared point like the oracle does."""
    B, S = 40, 12
    rng = np.random.default_rng(21)
    rec = m.scenarios.generate(B, 2, seed=14, weight_goal_1=20.0)
    rec[:, :, o2.RB:o2.RB + 6] = [0.08, 0.07, 0.09, 0.06, 0.08, 0.1]
    obst = random_obstacles(rng, B, 2, S, rec)
    fab = get_fab(fabs, 2)
    ref = oracle_actions(rec, obst)
    ok = np.isfinite(ref).all(axis=(1, 2))
    assert ok.sum() > 0.8 * B
    act = fab.action_host(rec, obst, dtype="f64")
    assert rel_ps(act, ref, ok) < F64_RTOL
    act32 = fab.action_host(rec, obst, dtype="f32")
    assert np.abs(act32 - ref)[ok].max() < 2e-3


def test_action_acc_mode_and_grasp_planner(fabs):
    """mode 'acc' returns qdd; has_collision_links=0 is the grasp planner (example_pandas_Jointspace.py:160-166)."""
    B = 20
    rng = np.random.default_rng(5)
    rec = m.scenarios.generate(B, 2, seed=12)
    obst = random_obstacles(rng, B, 2, 4, rec)
    for kw in (dict(mode=0), dict(has_collision_links=0)):
        fab = get_fab(fabs, 2, **kw)
        act = fab.action_host(rec, obst, dtype="f64")
        ref = oracle_actions(rec, obst, **kw)
        assert np.abs(act - ref).max() / np.abs(ref).max() < F64_RTOL


def test_cartesian_rollout_matches_oracle(fabs):
    B, N, S = 24, 20, 16
    rng = np.random.default_rng(9)
    rec = m.scenarios.generate(B, 2, seed=13, weight_goal_1=20.0)
    obst = random_obstacles(rng, B, 1, S, rec[:, :1] if rec.ndim == 3 else rec[:, None])[:, 0]
    fab = get_fab(fabs, 2)
    ocfg = o2.default_config(2)
    for robot in (0, 1):
        avg, qN, qdN = fab.rollout_cart_host(robot, rec[:, robot], obst, N, dtype="f64")
        n_cmp = 0
        for b in range(B):
            rq, rqd, ravg = o2.rollout_cartesian(ocfg, robot, rec[b, robot], obst[b, :, 0:3], obst[b, :, 3:6],
                                                 obst[b, :, 9], N)
            if not np.isfinite(rqd).all() or np.abs(rqd).max() > 10:
                continue
            assert np.abs(qdN[b] - rqd).max() / np.abs(rqd).max() < F64_RTOL
            assert np.abs(qN[b] - rq).max() < 1e-9 * np.abs(rq).max()
            assert abs(avg[b] - ravg) < 1e-9 * abs(ravg)
            n_cmp += 1
        assert n_cmp > B // 2


def test_kinematics_matches_oracle(fabs):
    R, B = 3, 17
    rec = m.scenarios.generate(B, R, seed=14)
    fab = get_fab(fabs, R)
    x, v, a = fab.kinematics_host(rec[..., 0:7], rec[..., 7:14])
    ocfg = o2.default_config(R)
    for b in range(B):
        for r in range(R):
            xx, vv, cc, _ = o2.kinematics(ocfg, r, rec[b, r, 0:7], rec[b, r, 7:14])
            assert np.abs(x[b, r] - xx).max() < 1e-12
            assert np.abs(v[b, r] - vv).max() < 1e-12
            assert np.abs(a[b, r] + cc).max() < 1e-11   # a = -(d(J qd)/dq) qd, utils.py:28


def test_device_api_equals_host_api(fabs):
    """SoA device-pointer entry (torch tensors) gives the host entry's bits; shard concatenation is bitwise."""
    import torch
    R, N, B = 3, 20, 100
    rec = m.scenarios.generate(B, R, seed=15)
    fab = get_fab(fabs, R)
    host = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    for dt, key in ((torch.float64, "f64"), (torch.float32, "f32")):
        d_rec = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=dt)
        qdN = torch.empty((R, N, 7, B), dtype=dt, device="cuda:0")
        xee = torch.empty((R, 3, B), dtype=dt, device="cuda:0")
        avg = fab.rollout_dev(d_rec, N, x_ee=xee, qdN=qdN)
        torch.cuda.synchronize()
        if key == "f64":
            assert np.array_equal(avg.cpu().numpy().T, host["avg_vel"])
            assert np.array_equal(qdN.permute(3, 0, 1, 2).cpu().numpy(), host["qdN"])
        # two shards == one batch, bitwise (the multi-GPU sharding property)
        h = B // 2
        a0 = fab.rollout_dev(d_rec[:, :, :h].contiguous(), N)
        a1 = fab.rollout_dev(d_rec[:, :, h:].contiguous(), N)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat([a0, a1], dim=1).view(torch.int32 if key == "f32" else torch.int64),
                           avg.view(torch.int32 if key == "f32" else torch.int64))


def test_full_size_batch_properties(fabs):
    """BASELINE config C5 shape (65536 x 3 Pandas x H50) in FP32: size-independent properties --
    batch-permutation equivariance and agreement of a strided sample with the oracle."""
    import torch
    R, N, B = 3, 50, 65536
    base = m.scenarios.generate(4096, R, seed=16)
    rec = np.tile(base, (B // 4096, 1, 1))
    fab = get_fab(fabs, R)
    d_rec = torch.from_numpy(to_soa(rec)).to("cuda:0", dtype=torch.float32)
    avg = fab.rollout_dev(d_rec, N)
    perm = torch.randperm(B, device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(0))
    avg_p = fab.rollout_dev(d_rec[:, :, perm].contiguous(), N)
    torch.cuda.synchronize()
    bits = lambda t: t.contiguous().view(torch.int32)            # bitwise (NaN-safe) comparison
    assert torch.equal(bits(avg[:, perm]), bits(avg_p))           # a scenario never reads another scenario
    assert torch.equal(bits(avg[:, :4096]), bits(avg[:, 4096:8192]))   # tiled copies give identical bits
    assert torch.isfinite(avg).float().mean() > 0.98
    idx = np.arange(0, 4096, 64)
    qN, qdN, ravg, xee, goal, ok = oracle_rollout(base[idx], R, N)
    got = avg[:, idx].cpu().numpy().T
    assert np.abs(got - ravg)[ok].max() < 1e-3


def test_host_api_pipelined_large_batch(fabs):
    """mrf_rollout_host chunks large batches (H2D of chunk c+1 overlaps the kernels of chunk c): results must equal the
    device-pointer entry bit for bit, including a batch that is not a multiple of the chunk / tile size."""
    import torch
    R, N = 3, 5
    for B in (8192, 10007):
        base = m.scenarios.generate(1024, R, seed=17)
        rec = np.tile(base, ((B + 1023) // 1024, 1, 1))[:B]
        rec[:, 0, 17] += np.arange(B) % 7 * 1e-3                     # make every scenario distinct
        fab = get_fab(fabs, R, estimate_goal=1)
        host = fab.rollout_host(rec.astype(np.float32), N, dtype="f32")
        d_rec = torch.from_numpy(to_soa(rec.astype(np.float32))).to("cuda:0")
        xee = torch.empty((R, 3, B), dtype=torch.float32, device="cuda:0")
        ge = torch.empty((3, B), dtype=torch.float32, device="cuda:0")
        avg = fab.rollout_dev(d_rec, N, x_ee=xee, goal_est=ge)
        torch.cuda.synchronize()
        assert np.array_equal(host["avg_vel"], avg.cpu().numpy().T, equal_nan=True)
        assert np.array_equal(host["x_ee"], xee.permute(2, 0, 1).cpu().numpy())
        assert np.array_equal(host["goal_est"], ge.cpu().numpy().T)


@pytest.mark.parametrize("R,n,vel_mode,sd", [(2, 4, 0, 1), (3, 4, 1, 1), (2, 1, 0, 0), (3, 2, 1, 1)])
def test_obstacle_staging_matches_oracle(fabs, R, n, vel_mode, sd):
    """mrf_obstacles: other robots' collision spheres with n_obst_per_link offsets, assembled per ego robot."""
    import torch
    from multi_robot_fabrics_b200.spheres import sphere_offsets
    B = 37
    rec = m.scenarios.generate(B, R, seed=61)
    fab = get_fab(fabs, R, static_or_dyn=sd)
    q = torch.from_numpy(np.ascontiguousarray(rec[:, :, 0:7].transpose(2, 1, 0))).to("cuda:0")
    qd = torch.from_numpy(np.ascontiguousarray(rec[:, :, 7:14].transpose(2, 1, 0))).to("cuda:0")
    sx = torch.empty((8 * n, 3, R, B), dtype=torch.float64, device="cuda:0")
    sv = torch.empty((8 * n, 3, R, B), dtype=torch.float64, device="cuda:0")
    obst = fab.obstacles_dev(q, qd, n_per_link=n, vel_mode=vel_mode, spheres_x=sx, spheres_v=sv)
    torch.cuda.synchronize()
    got = obst.permute(3, 2, 0, 1).cpu().numpy()                    # B,R,S,10
    ocfg = o2.default_config(R)
    off = sphere_offsets(n)
    for b in range(0, B, 6):
        ref = o2.obstacle_lists(ocfg, rec[b, :, 0:7], rec[b, :, 7:14], off, vel_mode=vel_mode, static_or_dyn=sd)
        for i in range(R):
            assert np.abs(got[b, i] - ref[i]).max() < 1e-12
        x, vo, vs = o2.spheres(ocfg, R - 1, rec[b, R - 1, 0:7], rec[b, R - 1, 7:14], off)
        assert np.abs(sx[:, :, R - 1, b].cpu().numpy() - x).max() < 1e-12
        assert np.abs(sv[:, :, R - 1, b].cpu().numpy() - (vs if vel_mode else vo) * sd).max() < 1e-12


@pytest.mark.parametrize("R,n", [(2, 4), (3, 4)])
def test_mrdf_control_step_configs_c2_c4(fabs, R, n):
    """BASELINE configs C2 / C4: one MRDF control step entirely on the GPU -- obstacle staging (n = 4 spheres per
    link: S = 32 for 2 Pandas, 64 for 3) feeding the executed action (weight_goal_1 = 20) -- against the oracle."""
    import torch
    from multi_robot_fabrics_b200.spheres import sphere_offsets
    B = 48
    rec = m.scenarios.generate(B, R, seed=62, weight_goal_1=20.0)
    fab = get_fab(fabs, R)
    d_rec = torch.from_numpy(to_soa(rec)).to("cuda:0")
    obst = fab.obstacles_dev(d_rec[0:7].contiguous(), d_rec[7:14].contiguous(), n_per_link=n, vel_mode=0)
    act = fab.action_dev(d_rec, obst)
    torch.cuda.synchronize()
    assert obst.shape[0] == 8 * n * (R - 1)
    got = act.permute(2, 1, 0).cpu().numpy()
    ocfg = o2.default_config(R)
    off = sphere_offsets(n)
    n_ok = 0
    for b in range(B):
        lists = o2.obstacle_lists(ocfg, rec[b, :, 0:7], rec[b, :, 7:14], off, vel_mode=0)
        for i in range(R):
            o = lists[i]
            try:
                ref = o2.action(ocfg, i, rec[b, i], o[:, 0:3], o[:, 3:6], o[:, 6:9], o[:, 9])
            except FloatingPointError:
                continue
            if not np.isfinite(ref).all() or np.abs(ref).max() > 5:
                continue
            n_ok += 1
            assert np.abs(got[b, i] - ref).max() < F64_RTOL * max(1.0, np.abs(ref).max())
    assert n_ok > 0.7 * B * R


def test_single_robot_rollout(fabs):
    """n_robots = 1: no inter-robot leaves, only plane / limit / attractor leaves (throughput kernel)."""
    R, N, B = 1, 20, 40
    rec = m.scenarios.generate(B, R, seed=81)
    fab = get_fab(fabs, R)
    out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N)
    assert ok.sum() > 0.9 * B
    assert rel_ps(out["qdN"], qdN, ok) < F64_RTOL
    assert np.abs(out["x_ee"] - xee).max() < 1e-12


@pytest.mark.parametrize("kernel", ["throughput", "cooperative"])
def test_four_robot_rollout_custom_mount(fabs, kernel):
    """n_robots = MRF_MAX_ROBOTS = 4 with a caller-supplied mount for the fourth arm (MrfConfig.mount)."""
    from multi_robot_fabrics_b200 import scenarios as sc
    R, N, B = 4, 10, 24
    old = sc.MOUNT_XYZ[3].copy()
    sc.MOUNT_XYZ[3] = [0.3, -0.6, 0.65]
    try:
        rec = sc.generate(B, R, seed=82)
        mounts = [sc.mount_matrix(r) for r in range(R)]
    finally:
        sc.MOUNT_XYZ[3] = old
    fab = Fabrics(R, device=0, mount=mounts)
    fab.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    fab.close()
    ocfg = o2.default_config(R)
    for r in range(R):
        for k, v in enumerate(mounts[r].reshape(16)):
            ocfg.mount[r][k] = v
    qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec, N)
    ok = np.isfinite(qdN).all(axis=(1, 2, 3)) & (np.abs(qdN).max(axis=(1, 2, 3)) < 3)
    assert ok.sum() > 0.8 * B
    assert rel_ps(out["qdN"], qdN, ok) < F64_RTOL
    assert rel_ps(out["avg_vel"], avg, ok) < F64_RTOL


def test_kernels_against_committed_golden_vectors(fabs):
    """The CUDA kernels against tests/golden/fabric_golden.npz directly (vectors derived by oracle O1, autodiff)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fabric_golden.npz"))
    # single actions (robots 0..2, S = 0..8, vel / acc mode, grasp planner, unequal radii, zero velocity)
    for c in range(int(g["n_act"])):
        p = f"act{c}_"
        kw = dict(mode=int(g[p + "mode"]), has_collision_links=0 if int(g[p + "grasp"]) else 1)
        fab = get_fab(fabs, 3, **kw)
        obst = g[p + "obst"]
        act = fab.action_host(g[p + "rec"][None, None], obst[None, None] if len(obst) else None,
                              robot_first=int(g[p + "robot"]), dtype="f64")[0, 0]
        assert np.abs(act - g[p + "action"]).max() < 1e-10 * max(1.0, np.abs(g[p + "action"]).max()), c
    # coupled rollouts: 2 Pandas, and 3 Pandas with unequal radii in dynamic and static mode, both kernels
    for coop in (0, 1 << 20):
        fab = get_fab(fabs, 2)
        fab.handle.set_coop_max_batch(coop)
        out = fab.rollout_host(g["ro_rec"][None], int(g["ro_N"]), dtype="f64", trajectories=True)
        fab.handle.set_coop_max_batch(512)
        assert np.abs(out["qN"][0] - g["ro_qN"]).max() < 1e-11 and np.abs(out["qdN"][0] - g["ro_qdN"]).max() < 1e-10
        assert np.abs(out["avg_vel"][0] - g["ro_avg"]).max() < 1e-10
        for sd in (1, 0):
            fab3 = Fabrics(3, device=0, static_or_dyn=sd, r_robots=g["ro3_rr"].tolist())
            fab3.handle.set_coop_max_batch(coop)
            out = fab3.rollout_host(g["ro3_rec"][None], int(g["ro3_N"]), dtype="f64", trajectories=True)
            fab3.close()
            assert np.abs(out["qdN"][0] - g[f"ro3_sd{sd}_qdN"]).max() < 1e-10
            assert np.abs(out["avg_vel"][0] - g[f"ro3_sd{sd}_avg"]).max() < 1e-10
    # Cartesian rollout and link kinematics
    fab = get_fab(fabs, 2)
    avg, qN, qdN = fab.rollout_cart_host(0, g["cart_rec"][None], g["cart_obst"][None], 2, dtype="f64")
    assert np.abs(qN[0] - g["cart_qN"]).max() < 1e-11 and np.abs(qdN[0] - g["cart_qdN"]).max() < 1e-10
    assert abs(avg[0] - float(g["cart_avg"])) < 1e-10
    fab3 = get_fab(fabs, 3)
    q = np.zeros((1, 3, 7)); qd = np.zeros((1, 3, 7))
    q[0, 1], qd[0, 1] = g["kin_rec"][0:7], g["kin_rec"][7:14]
    x, v, a = fab3.kinematics_host(q, qd)
    assert np.abs(x[0, 1] - g["kin_xva"][:, 0]).max() < 1e-13 and np.abs(v[0, 1] - g["kin_xva"][:, 1]).max() < 1e-13
    assert np.abs(a[0, 1] - g["kin_xva"][:, 2]).max() < 1e-12


def test_error_paths_fail_loudly(fabs):
    """Bad calls return an error code with a message instead of computing something else."""
    import ctypes as C
    import torch
    from multi_robot_fabrics_b200 import _lib
    rec = m.scenarios.generate(4, 2, seed=1)
    fab = get_fab(fabs, 2)
    with pytest.raises(_lib.MrfError, match="positive"):
        fab.rollout_host(rec, 0)                             # horizon must be positive
    L = _lib.lib()
    assert L.mrf_rollout_host_f64(fab.handle.ptr, None, 5, None, None, None, None, None, 4) == -1
    assert b"null" in L.mrf_last_error()
    d = torch.zeros((44, 2, 4), dtype=torch.float64, device="cuda:0")
    assert L.mrf_rollout_dev_f64(fab.handle.ptr, C.c_void_p(d.data_ptr()), 0, None, None, None, None, None, 4, None) == -1
    assert L.mrf_action_dev_f64(fab.handle.ptr, 1, 2, C.c_void_p(d.data_ptr()), 0, None, C.c_void_p(d.data_ptr()), 4, None) == -1
    with pytest.raises(_lib.MrfError):
        fab.rollout_dev(torch.zeros((44, 3, 4), dtype=torch.float64, device="cuda:0"), 5)     # wrong robot count
    with pytest.raises(_lib.MrfError):
        fab.rollout_dev(torch.zeros((44, 2, 4), dtype=torch.float16, device="cuda:0"), 5)     # unsupported dtype
    bad = _lib.default_config(2)
    bad.estimate_goal, bad.estimate_robot = 1, 5
    with pytest.raises(_lib.MrfError, match="estimate_robot"):
        m.Handle(bad)


def test_assumption_knobs_on_the_gpu(fabs):
    """jdot_sign / exec_scale / eps / jdot_ref_sign are MrfConfig fields: the kernels follow the oracle for the
    alternative conventions too (rollout and action, both rollout kernels)."""
    R, N, B = 2, 8, 24
    kn = dict(jdot_sign=1.0, exec_scale=0.5, eps=1e-5, jdot_ref_sign=1.0)
    rec = m.scenarios.generate(B, R, seed=91)
    ocfg = o2.default_config(R, **kn)
    qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec, N)
    ok = np.isfinite(qdN).all(axis=(1, 2, 3)) & (np.abs(qdN).max(axis=(1, 2, 3)) < 3)
    fab = Fabrics(R, device=0, **kn)
    for coop in (0, 1 << 20):
        fab.handle.set_coop_max_batch(coop)
        out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
        assert rel_ps(out["qdN"], qdN, ok) < F64_RTOL
    rng = np.random.default_rng(3)
    obst = random_obstacles(rng, B, R, 6, rec)
    act = fab.action_host(rec, obst, dtype="f64")
    ref = oracle_actions(rec, obst, **kn)
    okb = np.isfinite(ref).all(axis=(1, 2))
    assert rel_ps(act, ref, okb) < F64_RTOL
    fab.close()
    dflt = oracle_actions(rec, obst)
    assert np.abs(dflt - ref)[okb].max() > 1e-6


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_page_locked_host_records_are_read_in_place(built, dtype):
    """mrf_rollout_host with page-locked buffers (the kernel reads the records in record order straight from host
    memory) returns bit-identical results to the staged path with pageable buffers; ragged tail tile included."""
    import torch
    from multi_robot_fabrics_b200.api import Fabrics
    from multi_robot_fabrics_b200 import scenarios
    R, N, B = 3, 6, 8192 + 37
    base = scenarios.generate(512, R, seed=21)
    rec = np.tile(base, (B // 512 + 1, 1, 1))[:B].astype(np.float32 if dtype == "f32" else np.float64)
    fab = Fabrics(R, estimate_goal=1)
    ref = fab.rollout_host(rec, N, dtype=dtype)                                 # pageable -> staged pipeline
    tdt = torch.float32 if dtype == "f32" else torch.float64
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    out = {"avg_vel": pin(np.zeros((B, R), rec.dtype)), "x_ee": pin(np.zeros((B, R, 3), rec.dtype)),
           "goal_est": np.zeros((B, 3), rec.dtype)}                             # one result buffer left pageable
    n0 = fab.handle.launches
    got = fab.rollout_host(pin(rec), N, dtype=dtype, out=out)
    assert fab.handle.launches - n0 == 1                                  # one kernel, no transposes
    for k in ("avg_vel", "x_ee", "goal_est"):
        assert np.array_equal(got[k].view(np.uint8), ref[k].view(np.uint8)), k
    fab.close()


def test_degenerate_inputs_propagate_like_the_oracle(fabs):
    """The reference has no guards (SURVEY 8b 'error conventions'): a joint on its limit is 1/0 in the limit leaf, a
    joint beyond it or a penetrating sphere flips a leaf coordinate negative, a NaN input poisons the action.  The
    CUDA path must return NaN actions exactly where the oracle does and agree everywhere else.  (Coincidences that
    depend on the last bit of the forward kinematics -- goal exactly at the hand, sphere centre exactly on a link --
    are not testable across two implementations of the chain.)"""
    B, R, S = 8, 2, 3
    rng = np.random.default_rng(77)
    rec = m.scenarios.generate(B, R, seed=31, weight_goal_1=20.0)
    obst = random_obstacles(rng, B, R, S, rec)
    cfg = o2.default_config(R)
    for r in range(R):
        rec[1, r, 3] = float(cfg.limits[3][0])                # scenario 1: joint 4 on its lower limit -> 1/0
        rec[2, r, 5] = float(cfg.limits[5][1]) + 0.05         # scenario 2: joint 6 beyond its upper limit (x < 0)
        x, _, _, _ = o2.kinematics(cfg, r, rec[3, r, 0:7], rec[3, r, 7:14])
        obst[3, r, 0, 0:3] = x[4] + np.array([0.03, -0.02, 0.04])   # scenario 3: sphere overlapping link5 (x < 0)
        rec[4, r, 7:14] = 0.0                                 # scenario 4: robot at rest (sign switches at xdot = 0)
        rec[5, r, 2] = np.nan                                 # scenario 5: NaN joint position
        obst[6, r, 1, 3:6] = np.inf                           # scenario 6: infinite obstacle velocity
    fab = get_fab(fabs, R)
    ref = oracle_actions(rec, obst)
    for dtype, tol in (("f64", 1e-9), ("f32", 2e-3)):
        act = fab.action_host(rec, obst, dtype=dtype).astype(np.float64)
        assert np.array_equal(np.isnan(act), np.isnan(ref)), dtype
        okm = np.isfinite(ref)
        assert np.isfinite(act[okm]).all()
        if dtype == "f32":          # the overlapping-sphere rows are stiff (|action| ~ 1e3): FP32 is compared on the rest
            okm &= (np.abs(ref) < 10).all(axis=2, keepdims=True)
        assert np.abs(act - ref)[okm].max() / np.abs(ref[okm]).max() < tol, dtype
    assert np.isnan(ref[1]).all() and np.isnan(ref[5]).all() and np.isfinite(ref[[0, 3, 4]]).all()


def test_submit_wait_pipeline_equals_blocking_call(built):
    """mrf_rollout_host_submit / _wait (two batches in flight) return what the blocking entry returns, batch by batch."""
    import torch
    from multi_robot_fabrics_b200 import scenarios
    R, N, B = 3, 4, 4096 + 5
    fab = Fabrics(R, estimate_goal=1)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    recs = [pin(np.tile(scenarios.generate(256, R, seed=40 + i), (17, 1, 1))[:B].astype(np.float32)) for i in range(5)]
    outs = [{"avg_vel": pin(np.zeros((B, R), np.float32)), "x_ee": pin(np.zeros((B, R, 3), np.float32)),
             "goal_est": pin(np.zeros((B, 3), np.float32))} for _ in range(5)]
    for rec, out in zip(recs, outs):
        fab.rollout_host_submit(rec, N, out, dtype="f32")
    fab.rollout_host_wait(all=True)
    fab.handle.set_coop_max_batch(0)                       # the blocking reference through the same (throughput) kernel
    for rec, out in zip(recs, outs):
        ref = fab.rollout_host(np.array(rec), N, dtype="f32")
        for k in ref:
            assert np.array_equal(out[k].view(np.uint8), ref[k].view(np.uint8)), k
    with pytest.raises(m.MrfError):
        fab.rollout_host_submit(np.array(recs[0]), N, outs[0], dtype="f32")      # pageable records are refused
    # compact records: only q, qdot, x_goal_0, weight_goal_0 per scenario, the other fields once per submission
    assert all(np.all(r[:, :, 18:] == r[0:1, :, 18:]) for r in recs)
    outs_c = [{k: pin(np.zeros_like(v)) for k, v in o.items()} for o in outs]
    for rec, out in zip(recs, outs_c):
        fab.rollout_host_submit(pin(rec[:, :, :18]), N, out, dtype="f32", shared=rec[0])
    fab.rollout_host_wait(all=True)
    for out, ref in zip(outs_c, outs):
        for k in ref:
            assert np.array_equal(out[k].view(np.uint8), ref[k].view(np.uint8)), k
    fab.close()


@pytest.mark.parametrize("kernel", ["throughput", "cooperative"])
@pytest.mark.parametrize("links", [[[5], [5]], [[3, 6, 8], [1, 2, 4, 7]], [[1, 2, 3, 4, 5, 6, 7, 8], [2, 6]]])
def test_rollout_collision_link_subsets(fabs, kernel, links):
    """arbitrary collision_links_nr per robot (signature default [5],
    example_pandas_Jointspace.py:64).  The links in a robot's set are its ego leaves (link3..8) AND the spheres the other
    robots see; different radii on link5 / link6 exercise the shared-point logic."""
    R, N, B = 2, 10, 48
    rec = m.scenarios.generate(B, R, seed=31)
    rr = [[0.08, 0.06, 0.08, 0.08, 0.07, 0.09, 0.08, 0.08], [0.05, 0.05, 0.08, 0.1, 0.08, 0.08, 0.06, 0.08]]
    for r in range(R):
        rec[:, r, o2.RB:o2.RB + 6] = [0.08 if (l in links[r]) else 0.0 for l in range(3, 9)]
    rec[:, 0, o2.RB + 3] = 0.07 if 6 in links[0] else 0.0       # link6 radius differs from link5
    fab = Fabrics(R, device=0, collision_links=links, r_robots=rr)
    fab.handle.set_coop_max_batch(0 if kernel == "throughput" else 1 << 20)
    ocfg = o2.set_collision_links(o2.default_config(R), links)
    for r in range(R):
        for l in range(8):
            ocfg.r_robots[r][l] = rr[r][l]
    qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec, N)
    ok = np.isfinite(qdN).all(axis=(1, 2, 3)) & (np.abs(qdN).max(axis=(1, 2, 3)) < 3)
    assert ok.sum() > 0.8 * B
    out = fab.rollout_host(rec, N, dtype="f64", trajectories=True)
    assert rel_ps(out["qdN"], qdN, ok) < F64_RTOL and rel_ps(out["avg_vel"], avg, ok) < F64_RTOL
    out32 = fab.rollout_host(rec, N, dtype="f32", trajectories=True)
    assert np.quantile(np.abs(out32["qdN"] - qdN).max(axis=(1, 2, 3))[ok], 0.9) < 1e-4
    # the executed action with the same subsets (action kernel: ego-major FP64 and obstacle-major FP32 paths)
    rng = np.random.default_rng(4)
    obst = random_obstacles(rng, B, R, 9, rec)
    ref = np.full((B, R, 7), np.nan)
    for b in range(B):
        for r in range(R):
            o = obst[b, r]
            try:
                ref[b, r] = o2.action(ocfg, r, rec[b, r], o[:, 0:3], o[:, 3:6], o[:, 6:9], o[:, 9])
            except FloatingPointError:
                pass
    okb = np.isfinite(ref).all(axis=(1, 2))
    act = fab.action_host(rec, obst, dtype="f64")
    assert rel_ps(act, ref, okb) < F64_RTOL
    act32 = fab.action_host(rec, obst, dtype="f32")
    assert np.abs(act32 - ref)[okb].max() < 2e-3
    fab.close()


@pytest.mark.parametrize("R,S", [(2, 2), (3, 5)])
def test_rollout_with_static_spheres(fabs, R, S):
    """static spheres inside the coupled rollouts (x_obsts / radius_obsts of the rollout planners,
    forward_planner_Jointspace.py:319-322) through mrf_rollout_static_dev, against the oracle."""
    import torch
    N, B = 12, 40
    rec = m.scenarios.generate(B, R, seed=33)
    rng = np.random.default_rng(6)
    stat = np.zeros((B, R, S, 4))
    stat[..., 0:3] = rng.uniform([-0.4, -1.0, 1.5], [1.4, 1.0, 2.2], size=(B, R, S, 3))     # above the arms' workspace
    stat[:, :, 0, 0:3] = rng.uniform([0.0, -0.6, 0.9], [1.0, 0.6, 1.4], size=(B, R, 3))      # one inside it
    stat[..., 3] = rng.uniform(0.05, 0.12, size=(B, R, S))
    ocfg = o2.default_config(R)
    qdN, avg = np.zeros((B, R, N, 7)), np.zeros((B, R))
    for b in range(B):
        _, qdN[b], avg[b], _ = o2.rollout_jointspace_static(ocfg, rec[b], N, stat[b, :, :, 0:3], stat[b, :, :, 3])
    plain = o2.rollout_jointspace(ocfg, rec, N)[1]
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(qdN).all(axis=(1, 2, 3)) & (np.abs(qdN).max(axis=(1, 2, 3)) < 3)
    assert ok.sum() > 0.6 * B and np.abs(plain - qdN)[ok].max() > 1e-4       # the static spheres matter
    fab = get_fab(fabs, R)
    for dt, tol in ((torch.float64, None), (torch.float32, 1e-4)):
        dev = "cuda:0"
        d_rec = torch.from_numpy(to_soa(rec)).to(dev, dtype=dt)
        d_st = torch.from_numpy(np.ascontiguousarray(stat.transpose(2, 3, 1, 0))).to(dev, dtype=dt)
        q_o = torch.empty((R, N, 7, B), dtype=dt, device=dev)
        a_o = fab.rollout_static_dev(d_rec, d_st, N, qdN=q_o)
        torch.cuda.synchronize()
        got = q_o.permute(3, 0, 1, 2).double().cpu().numpy()
        if tol is None:
            assert rel_ps(got, qdN, ok) < F64_RTOL and rel_ps(a_o.T.cpu().numpy(), avg, ok) < F64_RTOL
        else:
            assert np.quantile(np.abs(got - qdN).max(axis=(1, 2, 3))[ok], 0.9) < tol
    with pytest.raises(m.MrfError):
        fab.rollout_static_dev(d_rec, torch.zeros((17, 4, R, B), dtype=dt, device=dev), N)     # > MRF_MAX_STATIC


def test_cartesian_rollout_acc_mode(fabs):
    """decoupled rollouts in fabrics_mode 'acc' (FabricsRollouts' constructor default,
    forward_planner_Cartesian.py:20,81-84): the action is qdd, pos += dt vel + dt^2/2 qdd, vel += dt qdd."""
    N, S, B = 10, 6, 32
    rng = np.random.default_rng(8)
    rec = m.scenarios.generate(B, 2, seed=35, weight_goal_1=20.0)[:, :1]
    obst = random_obstacles(rng, B, 1, S, rec)
    fab = Fabrics(2, device=0, mode=0)
    ocfg = o2.default_config(2, mode=0)
    avg, qN, qdN = fab.rollout_cart_host(0, rec[:, 0], obst[:, 0], N, dtype="f64")
    n_cmp = 0
    for b in range(B):
        rq, rqd, ravg = o2.rollout_cartesian(ocfg, 0, rec[b, 0], obst[b, 0, :, 0:3], obst[b, 0, :, 3:6], obst[b, 0, :, 9], N)
        if not np.isfinite(rqd).all() or np.abs(rqd).max() > 3:
            continue
        assert np.abs(qdN[b] - rqd).max() < 1e-9 * max(1.0, np.abs(rqd).max()) and np.abs(qN[b] - rq).max() < 1e-9
        assert abs(avg[b] - ravg) < 1e-9 * max(1.0, abs(ravg))
        n_cmp += 1
    assert n_cmp > B // 2
    fab.close()



def test_full_size_c5_properties(fabs):
    """BASELINE config C5 at full size (65 536 scenarios x 3 Pandas x horizon 50) through size-independent properties:
    idempotence (the same batch twice -> bitwise equal), sharding (the halves rolled separately concatenate bitwise to the
    whole), agreement of the FP32 kernel with the FP64 kernel in distribution, and the oracle on a sample of the batch."""
    import torch
    R, N, B = 3, 50, 65536
    rec = m.scenarios.generate(B, R, seed=2025).astype(np.float32)
    fab = get_fab(fabs, R, estimate_goal=1)
    dev = "cuda:0"
    d = torch.from_numpy(to_soa(rec)).to(dev)
    a1 = fab.rollout_dev(d, N)
    a2 = fab.rollout_dev(d, N)
    h = B // 2
    lo = fab.rollout_dev(d[:, :, :h].contiguous(), N)
    hi = fab.rollout_dev(d[:, :, h:].contiguous(), N)
    a64 = fab.rollout_dev(d.double(), N)
    torch.cuda.synchronize()
    assert torch.equal(a1.view(torch.int32), a2.view(torch.int32))
    assert torch.equal(torch.cat([lo, hi], dim=1).view(torch.int32), a1.view(torch.int32))
    # over 50 explicit-Euler steps ~1 % of random scenarios blow up to non-finite values -- in BOTH precisions (the reference
    # would, too); FP32 must not lose more than a few per mille beyond those
    f32ok = torch.isfinite(a1).all(dim=0).cpu().numpy()
    f64ok = torch.isfinite(a64).all(dim=0).cpu().numpy()
    assert f64ok.mean() > 0.97 and (f32ok | ~f64ok).mean() > 0.997
    both = f32ok & f64ok
    e = (a1.double() - a64).abs().amax(dim=0).cpu().numpy()[both]
    assert np.median(e) < 1e-6 and np.quantile(e, 0.99) < 1e-3
    idx = np.arange(0, B, B // 256)
    ref = o2.rollout_rfcv(o2.default_config(R), rec[idx].astype(np.float64), N)
    got = a64[:, torch.from_numpy(idx).to(dev)].T.cpu().numpy()
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(ref["avg_vel"]).all(axis=1) & (ref["avg_vel"].max(axis=1) < 2.0)
    assert ok.sum() > 200 and rel_ps(got, ref["avg_vel"], ok) < F64_RTOL
