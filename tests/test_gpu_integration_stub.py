"""The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add) works as written: it binds the
C-ABI without this repository's Python package and reproduces the package's results."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class MrfConfig(C.Structure):          # field order of include/mrf_b200.h
    _fields_ = [("struct_size", C.c_int32), ("n_robots", C.c_int32), ("mode", C.c_int32), ("static_or_dyn", C.c_int32),
                ("has_collision_links", C.c_int32), ("estimate_goal", C.c_int32), ("estimate_robot", C.c_int32),
                ("reserved0", C.c_int32), ("estimate_horizon", C.c_double), ("dt", C.c_double), ("eps", C.c_double),
                ("jdot_sign", C.c_double), ("jdot_ref_sign", C.c_double), ("exec_scale", C.c_double),
                ("mount", (C.c_double * 16) * 4), ("limits", (C.c_double * 2) * 7), ("r_robots", (C.c_double * 8) * 4),
                ("dl_avg_vel_constant", C.c_double), ("dl_dist_constant", C.c_double),
                ("dl_goal_weight_follower", C.c_double), ("dl_goal_weight_leader", C.c_double),
                ("dl_nr_goal_scale", C.c_double), ("dl_dist_endeff", C.c_double), ("dl_backoff", C.c_double),
                ("dl_time_wait", C.c_int32), ("dl_time_gate", C.c_int32), ("collision_link_mask", C.c_int32 * 4)]


def test_integration_md_stub(built):
    import torch
    import multi_robot_fabrics_b200 as m
    from multi_robot_fabrics_b200.api import Fabrics
    L = C.CDLL(os.path.join(ROOT, "multi-robot-fabrics_b200", "libmrf_b200.so"))
    L.mrf_last_error.restype = C.c_char_p
    cfg, h = MrfConfig(), C.c_void_p()
    assert L.mrf_config_default(C.byref(cfg), 3) == 0
    assert L.mrf_create(C.byref(cfg), 0, C.byref(h)) == 0, L.mrf_last_error().decode()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    B, R, N = 300, 3, 8
    rec = m.scenarios.generate(B, R, seed=3)
    avg, xee, goal = np.empty((B, R)), np.empty((B, R, 3)), np.empty((B, 3))
    rc = L.mrf_rollout_host_f64(h, p(np.ascontiguousarray(rec)), C.c_int(N), p(avg), p(xee), p(goal), None, None, C.c_int64(B))
    assert rc == 0, L.mrf_last_error().decode()
    ref = Fabrics(R).rollout_host(rec, N, dtype="f64")
    assert np.array_equal(avg.view(np.uint8), ref["avg_vel"].view(np.uint8))
    assert np.array_equal(xee.view(np.uint8), ref["x_ee"].view(np.uint8))
    # sweep stub: compact page-locked records, two batches in flight
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory().numpy()
    Bs = 8192
    big = np.tile(rec, (Bs // B + 1, 1, 1))[:Bs]
    shared = np.ascontiguousarray(big[0], dtype=np.float32)
    batches = [(pin(np.roll(big, k, axis=0)[:, :, :18]), pin(np.zeros((Bs, R)))) for k in range(3)]
    for rec_var, out in batches:
        rc = L.mrf_rollout_host_submit_compact_f32(h, p(rec_var), p(shared), C.c_int(N), p(out), None, None, C.c_int64(Bs))
        assert rc == 0, L.mrf_last_error().decode()
    assert L.mrf_rollout_host_wait(h, C.c_int(1)) == 0
    fab = Fabrics(R)
    fab.handle.set_coop_max_batch(0)
    for k, (rec_var, out) in enumerate(batches):
        want = fab.rollout_host(np.roll(big, k, axis=0).astype(np.float32), N, dtype="f32")["avg_vel"]
        assert np.array_equal(out.view(np.uint8), want.view(np.uint8)), k
    # RF-CV control-step stub: rollout + vel_avg_tot + deadlock_checking for a batch, result (R+1, B) back
    cfg2, h2 = MrfConfig(), C.c_void_p()
    assert L.mrf_config_default(C.byref(cfg2), 3) == 0
    cfg2.estimate_goal = 1                                   # ESTIMATE_GOAL: RF-CV
    assert L.mrf_create(C.byref(cfg2), 0, C.byref(h2)) == 0, L.mrf_last_error().decode()
    results = [pin(np.zeros((R + 1, Bs))) for _ in range(3)]
    for (rec_var, _), res in zip(batches, results):
        rc = L.mrf_rfcv_host_submit_f32(h2, p(rec_var), p(shared), C.c_int(N), C.c_int(100), p(res), None, C.c_int64(Bs))
        assert rc == 0, L.mrf_last_error().decode()
    assert L.mrf_rollout_host_wait(h2, C.c_int(1)) == 0
    fab2 = Fabrics(R, estimate_goal=1)
    for k, res in enumerate(results):
        want = np.zeros((R + 1, Bs), dtype=np.float32)
        hr = pin(np.roll(big, k, axis=0))
        fab2.rfcv_host_submit(hr, N, pin_res := pin(want), time_step=100)
        fab2.rollout_host_wait(all=True)
        assert np.array_equal(res.view(np.uint8), pin_res.view(np.uint8)), k
        assert set(np.unique(res[R])) <= {0.0, 1.0}
    assert L.mrf_destroy(h2) == 0
    assert L.mrf_destroy(h) == 0
