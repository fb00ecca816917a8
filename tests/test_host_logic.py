"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, fails loudly without a
GPU, packs the reference's kwargs correctly, and the device code (compiled for the host by tests/emul, a debugging
aid) reproduces the oracle -- so kernel logic errors show up before a GPU is involved."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import multi_robot_fabrics_b200 as m
from multi_robot_fabrics_b200 import _lib
from oracle import o2

from helpers import oracle_actions, oracle_rollout, random_obstacles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(built):
    L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libmrf_emul.so"))
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def emul_rollout(L, cfg, rec, N, prec="f64"):
    B, R = rec.shape[:2]
    o = dict(avg=np.zeros((B, R)), xee=np.zeros((B, R, 3)), goal=np.zeros((B, 3)), qN=np.zeros((B, R, N, 7)),
             qdN=np.zeros((B, R, N, 7)))
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    getattr(L, f"emul_rollout_{prec}")(C.byref(cfg), _dp(rec), N, _dp(o["avg"]), _dp(o["xee"]), _dp(o["goal"]),
                                       _dp(o["qN"]), _dp(o["qdN"]), C.c_longlong(B))
    return o


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "mrf_b200.h")).read()
    declared = set(re.findall(r"\b(mrf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.mrf_version() >= 100


def test_config_default_is_the_reference_setup(built):
    cfg = _lib.default_config(3)
    assert cfg.struct_size == C.sizeof(_lib.MrfConfig)
    assert (cfg.n_robots, cfg.mode, cfg.static_or_dyn, cfg.dt, cfg.eps) == (3, 1, 1, 0.01, 1e-6)
    assert [cfg.mount[r][3] for r in range(3)] == [0.0, 1.0, 0.7] and cfg.mount[2][7] == 0.6    # parameters_manipulators.py:98-102
    assert abs(cfg.mount[1][0] + 1.0) < 1e-15 and cfg.mount[0][0] == 1.0                         # yaw pi / 0
    assert cfg.limits[3][0] == -3.0718 and cfg.limits[5][1] == 3.7525
    assert (cfg.dl_avg_vel_constant, cfg.dl_dist_endeff, cfg.dl_time_wait, cfg.dl_time_gate) == (0.16, 0.35, 300, 10)
    o = o2.default_config(3)
    for r in range(3):
        assert np.allclose(np.array(cfg.mount[r][:]), np.array(o.mount[r][:]))


def test_no_gpu_fails_loudly(built):
    """No CPU fallback: without a device the handle cannot be created and every planner constructor raises."""
    if _lib.lib().mrf_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.MrfError, match="no CUDA device"):
        m.Handle(_lib.default_config(2))
    from multi_robot_fabrics_b200 import planner
    with pytest.raises(_lib.MrfError):
        planner.set_planner_panda(7, 0, 8, [1, 2, 3, 4, 5, 6, 7, 8], {}, {"mount_positions": [np.zeros(3)]}, 0)


def test_bad_arguments_are_rejected(built):
    L = _lib.lib()
    cfg = _lib.default_config(2)
    assert L.mrf_config_default(C.byref(cfg), 9) == _lib.lib().mrf_config_default(C.byref(cfg), 0) != 0
    bad = _lib.default_config(2)
    bad.struct_size = 12
    h = C.c_void_p()
    assert L.mrf_create(C.byref(bad), 0, C.byref(h)) == -1
    assert b"size mismatch" in L.mrf_last_error()
    with pytest.raises(_lib.MrfError):
        _lib.default_config(2, not_a_field=1)


@pytest.mark.parametrize("R,N,est", [(2, 20, 0), (3, 12, 1), (2, 5, 2)])
def test_emulated_kernel_rollout_matches_oracle(emul, R, N, est):
    B = 40
    rec = m.scenarios.generate(B, R, seed=5 + R)
    cfg = _lib.default_config(R, estimate_goal=est)
    qN, qdN, avg, xee, goal, ok = oracle_rollout(rec, R, N, estimate_goal=est)
    o = emul_rollout(emul, cfg, rec, N, "f64")
    assert ok.sum() > 0.9 * B
    assert np.abs(o["qdN"] - qdN)[ok].max() / np.abs(qdN[ok]).max() < 1e-9
    assert np.abs(o["avg"] - avg)[ok].max() < 1e-9
    assert np.abs(o["xee"] - xee).max() < 1e-13 and np.abs(o["goal"] - goal).max() < 1e-13
    o32 = emul_rollout(emul, cfg, rec, N, "f32")
    assert np.abs(o32["qdN"] - qdN)[ok].max() < 2e-3


def test_emulated_kernel_nonuniform_radii_and_static(emul):
    """Sphere-table merging (link1==link2, link5==link6) must fall back to separate entries when radii differ."""
    R, N, B = 2, 6, 16
    rec = m.scenarios.generate(B, R, seed=21)
    rec[:, :, o2.RB:o2.RB + 6] = [0.08, 0.07, 0.09, 0.06, 0.08, 0.1]       # link5 != link6 body radius
    rr = [[0.08, 0.06, 0.08, 0.08, 0.07, 0.09, 0.08, 0.08], [0.05, 0.05, 0.08, 0.1, 0.08, 0.08, 0.06, 0.08]]
    for sd in (1, 0):
        cfg = _lib.default_config(R, r_robots=rr, static_or_dyn=sd)
        ocfg = o2.default_config(R, static_or_dyn=sd)
        for r in range(R):
            for l in range(8):
                ocfg.r_robots[r][l] = rr[r][l]
        qN, qdN, avg, _ = o2.rollout_jointspace(ocfg, rec, N)
        o = emul_rollout(emul, cfg, rec, N)
        ok = np.isfinite(qdN).all(axis=(1, 2, 3))
        assert ok.sum() > 10
        assert np.abs(o["qdN"] - qdN)[ok].max() / np.abs(qdN[ok]).max() < 1e-9


@pytest.mark.parametrize("kw", [dict(), dict(mode=0), dict(has_collision_links=0)])
def test_emulated_kernel_action_matches_oracle(emul, kw):
    B, S = 24, 7
    rng = np.random.default_rng(2)
    rec = m.scenarios.generate(B, 2, seed=31, weight_goal_1=20.0)
    obst = random_obstacles(rng, B, 2, S, rec)
    cfg = _lib.default_config(2, **kw)
    ref = oracle_actions(rec, obst, **kw)
    for robot in (0, 1):
        out = np.zeros((B, 7))
        emul.emul_action_f64(C.byref(cfg), robot, _dp(np.ascontiguousarray(rec[:, robot])), S,
                             _dp(np.ascontiguousarray(obst[:, robot])), _dp(out), C.c_longlong(B))
        assert np.abs(out - ref[:, robot]).max() / np.abs(ref[:, robot]).max() < 1e-9


def test_emulated_kernel_cartesian_rollout(emul):
    B, S, N = 6, 5, 10
    rng = np.random.default_rng(4)
    rec = m.scenarios.generate(B, 2, seed=41, weight_goal_1=20.0)[:, 0]
    obst = random_obstacles(rng, B, 1, S, rec[:, :1] if rec.ndim == 3 else rec[:, None])[:, 0]
    cfg, ocfg = _lib.default_config(2), o2.default_config(2)
    avg, qN, qdN = np.zeros(B), np.zeros((B, N, 7)), np.zeros((B, N, 7))
    emul.emul_cart_f64(C.byref(cfg), 0, _dp(np.ascontiguousarray(rec)), S, _dp(np.ascontiguousarray(obst)), N, _dp(avg),
                       _dp(qN), _dp(qdN), C.c_longlong(B))
    for b in range(B):
        rq, rqd, ravg = o2.rollout_cartesian(ocfg, 0, rec[b], obst[b, :, 0:3], obst[b, :, 3:6], obst[b, :, 9], N)
        assert np.abs(qdN[b] - rqd).max() / np.abs(rqd).max() < 1e-9 and abs(avg[b] - ravg) < 1e-9


def test_compute_action_kwarg_packing(built):
    """The kwargs of examples/example_pandas_Jointspace.py:421-439 land in the right record / obstacle slots."""
    from multi_robot_fabrics_b200.planner import PandaFabricPlanner
    pl = object.__new__(PandaFabricPlanner)             # packing only: no device needed
    pl.nr_obst, pl.nr_obst_dyn, pl.collision_links_nr = 0, 3, [3, 4, 5, 6, 7, 8]
    xs = [np.array([0.1 * i, 0.2, 1.0]) for i in range(3)]
    kw = dict(q=np.arange(7) * 0.1, qdot=np.arange(7) * -0.1, x_goal_0=np.array([0.4, 0.5, 0.6]), weight_goal_0=3,
              angle_goal_1=np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0]]), x_goal_1=np.array([0.107, 0.0, 0.0]),
              weight_goal_1=20.0, x_goal_2=np.array([np.pi / 4]), weight_goal_2=1.0, x_obsts=xs, radius_obsts=[0.08] * 3,
              constraint_0=np.array([0, 0, 1, -0.65]), radius_body_panda_links={str(l): np.array(0.08) for l in range(3, 9)},
              radius_body_panda_hand=np.array([0.08]), x_obsts_dynamic=xs, xdot_obsts_dynamic=[np.ones(3) * 0.1] * 3,
              xddot_obsts_dynamic=[np.zeros(3)] * 3, radius_obsts_dynamic=[0.08, 0.09, 0.1])
    rec, obst = pl._record_and_obstacles(kw)
    assert np.allclose(rec[_lib.Q:_lib.Q + 7], np.arange(7) * 0.1) and rec[_lib.W0] == 3 and rec[_lib.W1] == 20
    assert np.allclose(rec[_lib.ANG:_lib.ANG + 9], [0, 0, -1, 0, 1, 0, 1, 0, 0]) and rec[_lib.CON + 3] == -0.65
    assert np.allclose(rec[_lib.RB:_lib.RB + 6], 0.08) and obst.shape == (3, 10)
    assert np.allclose(obst[1, 0:3], xs[1]) and np.allclose(obst[:, 3:6], 0.1) and np.allclose(obst[:, 9], [0.08, 0.09, 0.1])
    kw2 = {k: v for k, v in kw.items() if "dynamic" not in k}
    for i in range(3):                                   # per-index spellings (example_pointmasses_dynamic.py:199-211)
        kw2[f"x_obst_dynamic_{i}"], kw2[f"xdot_obst_dynamic_{i}"] = xs[i], np.ones(3) * 0.1
        kw2[f"xddot_obst_dynamic_{i}"], kw2[f"radius_obst_dynamic_{i}"] = np.zeros(3), [0.08, 0.09, 0.1][i]
    rec2, obst2 = pl._record_and_obstacles(kw2)
    assert np.array_equal(rec, rec2) and np.array_equal(obst, obst2)


def test_scenarios_are_deterministic_and_clear(built):
    a, b = m.scenarios.generate(300, 3, seed=1), m.scenarios.generate(300, 3, seed=1)
    assert np.array_equal(a, b) and not np.array_equal(a, m.scenarios.generate(300, 3, seed=2))
    assert m.scenarios.clearance(a).min() >= 0.25
    x, *_ = o2.kinematics(o2.default_config(3), 2, a[0, 2, 0:7], a[0, 2, 7:14])
    assert np.abs(m.scenarios.link_positions(a[0, 2, 0:7], m.scenarios.mount_matrix(2)) - x).max() < 1e-12


def test_pick_and_place_layout_matches_oracle_kinematics(built):
    """The synthetic pick-and-place task: start goals are the hands' positions (oracle FK), blocks lie `drop` below."""
    rec = m.scenarios.generate(6, 3, seed=4)
    blocks, start = m.scenarios.pick_and_place_layout(rec, n_blocks=2, seed=1, spread=0.1, drop=0.2)
    assert blocks.shape == (6, 2, 3, 3) and start.shape == (6, 3, 3)
    cfg = o2.default_config(3)
    for b in range(6):
        for r in range(3):
            x = o2.kinematics(cfg, r, rec[b, r, 0:7], rec[b, r, 7:14])[0][7]
            assert np.abs(start[b, r] - x).max() < 1e-12
    assert np.allclose(blocks[..., 2], start[:, None, :, 2] - 0.2) and np.abs(blocks[..., :2] - start[:, None, :, :2]).max() <= 0.1
