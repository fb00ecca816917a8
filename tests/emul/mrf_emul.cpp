// Host emulation of the CUDA thread code -- DEBUGGING AID FOR THE CPU TEST SUITE ONLY.
// Compiles multi-robot-fabrics_b200/csrc/mrf_device.cuh for the host and walks a CTA tile thread by thread
// (Phase A for every thread, barrier, Phase B for every thread), so the logic of the kernels can be checked
// against the oracle in a container without a GPU.  It is NOT part of the product: libmrf_b200.so does not
// contain it and no product path calls it.  The GPU parity tests (-m gpu) run the real kernels.
#include <cstring>
#include <vector>

#include "../../multi-robot-fabrics_b200/csrc/mrf_devcfg.h"

using namespace mrf;

template <typename T>
static int rollout(const MrfConfig* mc, const double* rec, int N, double* avg, double* x_ee, double* goal_est,
                   double* qN, double* qdN, long long B) {
    DevCfg<T> cfg;
    fill_devcfg(*mc, cfg);
    const int R = cfg.n_robots, NT = kTile * R;
    std::vector<T> kin((size_t)kKinRows<T> * NT), prm((size_t)P_N * NT), q((size_t)NT * 7), qd((size_t)NT * 7), acc(NT);
    std::vector<Chain<T>> ch(NT);
    for (long long t0 = 0; t0 < B; t0 += kTile) {
        for (int tid = 0; tid < NT; ++tid) {
            int lane = tid % kTile, r = tid / kTile;
            long long b = t0 + lane;
            long long bb = b < B ? b : B - 1;
            auto ld = [&](int f) { return (T)rec[(bb * R + r) * MRF_REC + f]; };
            for (int i = 0; i < 7; ++i) {
                q[tid * 7 + i] = ld(MRF_Q + i);
                qd[tid * 7 + i] = ld(MRF_QD + i);
            }
            load_params<T>(ld, prm.data(), NT, tid);
            acc[tid] = 0;
            chain_forward(cfg, r, &q[tid * 7], &qd[tid * 7], ch[tid], kin.data(), NT, tid);
            V3<T> p8 = kin_load(kin.data(), NT, tid, 4, 0);
            if (x_ee && b < B) {
                x_ee[(b * R + r) * 3 + 0] = p8.x; x_ee[(b * R + r) * 3 + 1] = p8.y; x_ee[(b * R + r) * 3 + 2] = p8.z;
            }
            if (r == cfg.estimate_robot) {
                V3<T> g = mk(prm[(P_G0 + 0) * NT + tid], prm[(P_G0 + 1) * NT + tid], prm[(P_G0 + 2) * NT + tid]);
                if (cfg.estimate_goal) {
                    V3<T> l1 = mk(cfg.link1[r][0], cfg.link1[r][1], cfg.link1[r][2]);
                    V3<T> v = cfg.estimate_goal == 1 ? cross(ch[tid].z[0], p8 - l1) : kin_load(kin.data(), NT, tid, 4, 3);
                    g = p8 + v * cfg.est_h;
                    prm[(P_G0 + 0) * NT + tid] = g.x; prm[(P_G0 + 1) * NT + tid] = g.y; prm[(P_G0 + 2) * NT + tid] = g.z;
                }
                if (goal_est && b < B) { goal_est[b * 3] = g.x; goal_est[b * 3 + 1] = g.y; goal_est[b * 3 + 2] = g.z; }
            }
        }
        for (int k = 0; k < N; ++k) {
            for (int tid = 0; tid < NT; ++tid) {
                int r = tid / kTile;
                for (int i = 0; i < 7; ++i) q[tid * 7 + i] += cfg.dt * qd[tid * 7 + i];
                chain_forward(cfg, r, &q[tid * 7], &qd[tid * 7], ch[tid], kin.data(), NT, tid);
            }
            for (int tid = 0; tid < NT; ++tid) {
                int lane = tid % kTile, r = tid / kTile;
                long long b = t0 + lane;
                const T vref = cfg.static_or_dyn ? T(1) : T(0), aref = cfg.static_or_dyn ? cfg.sref : T(0);
                T act[7];
                if (cfg.uniform_obst && R == 3) {
                    SmemSrcUniform<T, 3> src{kin.data(), lane, r, vref, aref, cfg.r_obst};
                    fabric_action(cfg, r, &q[tid * 7], &qd[tid * 7], ch[tid], kin.data(), prm.data(), NT, tid, src, act);
                } else if (cfg.uniform_obst && R == 2) {
                    SmemSrcUniform<T, 2> src{kin.data(), lane, r, vref, aref, cfg.r_obst};
                    fabric_action(cfg, r, &q[tid * 7], &qd[tid * 7], ch[tid], kin.data(), prm.data(), NT, tid, src, act);
                } else {
                    SmemSrc<T> src{cfg, kin.data(), NT, lane, r, vref, aref};
                    fabric_action(cfg, r, &q[tid * 7], &qd[tid * 7], ch[tid], kin.data(), prm.data(), NT, tid, src, act);
                }
                for (int i = 0; i < 7; ++i) {
                    qd[tid * 7 + i] = act[i];
                    acc[tid] += act[i] * act[i];
                    if (b < B) {
                        if (qN) qN[(((b * R + r) * N) + k) * 7 + i] = q[tid * 7 + i];
                        if (qdN) qdN[(((b * R + r) * N) + k) * 7 + i] = act[i];
                    }
                }
            }
        }
        for (int tid = 0; tid < NT; ++tid) {
            int lane = tid % kTile, r = tid / kTile;
            long long b = t0 + lane;
            if (avg && b < B) avg[b * R + r] = acc[tid] / (T(N) * T(7));
        }
    }
    return 0;
}

template <typename T, bool CART>
static int action(const MrfConfig* mc, int robot, const double* rec, int S, const double* obst, int N, double* out,
                  double* qN, double* qdN, long long B) {
    DevCfg<T> cfg;
    fill_devcfg(*mc, cfg);
    const int NT = 1;
    std::vector<T> kin(kKinRows<T>), prm(P_N), ob((size_t)S * MRF_OBST + 1);
    for (long long b = 0; b < B; ++b) {
        auto ld = [&](int f) { return (T)rec[b * MRF_REC + f]; };
        T q[7], qd[7];
        for (int i = 0; i < 7; ++i) { q[i] = ld(MRF_Q + i); qd[i] = ld(MRF_QD + i); }
        load_params<T>(ld, prm.data(), NT, 0);
        for (int o = 0; o < S * MRF_OBST; ++o) ob[o] = (T)obst[b * S * MRF_OBST + o]; // [o][c] with stride 1
        Chain<T> ch;
        GlobalSrc<T, CART> src{ob.data(), 1, 0, S, T(0), T(1), T(1), nullptr, 1, 0};
        T act[7];
        if (!CART) {
            chain_forward(cfg, robot, q, qd, ch, kin.data(), NT, 0);
            fabric_action(cfg, robot, q, qd, ch, kin.data(), prm.data(), NT, 0, src, act);
            for (int i = 0; i < 7; ++i) out[b * 7 + i] = act[i];
        } else {
            T acc = 0;
            for (int k = 0; k < N; ++k) {
                src.tk = T(k) * cfg.dt;
                chain_forward(cfg, robot, q, qd, ch, kin.data(), NT, 0);
                fabric_action(cfg, robot, q, qd, ch, kin.data(), prm.data(), NT, 0, src, act);
                for (int i = 0; i < 7; ++i) {
                    qd[i] = act[i]; q[i] += cfg.dt * act[i]; acc += act[i] * act[i];
                    if (qN) qN[(b * N + k) * 7 + i] = q[i];
                    if (qdN) qdN[(b * N + k) * 7 + i] = qd[i];
                }
            }
            if (out) out[b] = acc / (T(N) * T(7));
        }
    }
    return 0;
}

extern "C" {
int emul_rollout_f64(const MrfConfig* c, const double* rec, int N, double* avg, double* x_ee, double* goal_est, double* qN, double* qdN, long long B) { return rollout<double>(c, rec, N, avg, x_ee, goal_est, qN, qdN, B); }
int emul_rollout_f32(const MrfConfig* c, const double* rec, int N, double* avg, double* x_ee, double* goal_est, double* qN, double* qdN, long long B) { return rollout<float>(c, rec, N, avg, x_ee, goal_est, qN, qdN, B); }
int emul_action_f64(const MrfConfig* c, int robot, const double* rec, int S, const double* obst, double* out, long long B) { return action<double, false>(c, robot, rec, S, obst, 0, out, nullptr, nullptr, B); }
int emul_action_f32(const MrfConfig* c, int robot, const double* rec, int S, const double* obst, double* out, long long B) { return action<float, false>(c, robot, rec, S, obst, 0, out, nullptr, nullptr, B); }
int emul_cart_f64(const MrfConfig* c, int robot, const double* rec, int S, const double* obst, int N, double* avg, double* qN, double* qdN, long long B) { return action<double, true>(c, robot, rec, S, obst, N, avg, qN, qdN, B); }
int emul_cart_f32(const MrfConfig* c, int robot, const double* rec, int S, const double* obst, int N, double* avg, double* qN, double* qdN, long long B) { return action<float, true>(c, robot, rec, S, obst, N, avg, qN, qdN, B); }
}
