// mrf_coop.cuh -- low-latency cooperative rollout: one CTA per scenario, one WARP per robot.
//
// The throughput kernel (rollout_kernel) gives each (scenario, robot) one thread, which is the right shape for
// thousands of scenarios but leaves a single rollout latency-bound (~9 us per horizon step in one thread).  Here the
// 32 lanes of a robot's warp split one fabric action:
//   * every lane runs the (cheap, serial) chain and the 14 limit leaves redundantly -- no intra-robot exchange;
//   * lane = ego point e (5) x other-robot point pt (6): one sphere leaf per lane and other robot, plane leaves on
//     lanes 0..4; a strided shuffle reduction over pt leaves the task-space metric A_e and force b_e on lanes 0..4;
//   * lanes 0..4 pull their ego point back (J^T A J, J^T b); lanes 6 and 7 pull back the two task-space attractors
//     with the same code (sub-goal 0 as a point load at link8, sub-goal 1 as A = m R^T R at the virtual point
//     hand - link7); an 8-lane butterfly sums the 27 entries of the geometry part and of the attractor part;
//   * even lanes factor M_G, odd lanes M_F (same instruction stream, different data), two shuffles broadcast the
//     solutions and every lane finishes the damper redundantly.
// Robots exchange their link points through shared memory exactly as in rollout_kernel (two barriers per step).
#pragma once

#include "mrf_device.cuh"

namespace mrf {

template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <typename T> __device__ __forceinline__ T shfl_t(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// Phase B of one robot, executed by its warp.  kin / prm are indexed [k * R + r] (one column per robot).
template <typename T, int R>
__device__ __forceinline__ void fabric_action_coop(const DevCfg<T>& cfg, int r, int lane, const T* q, const T* qd,
                                                   const Chain<T>& ch, const T* kin, const T* prm, T vref, T aref,
                                                   T* act) {
    constexpr int NT = R;
    const int tid = r;
    const T sigma = cfg.sigma;
    const T e_ = cfg.eps;
    // ---- joint-limit leaves, redundantly on every lane: diagonal metric, force, energy numerator ----
    T Gd[kDof], Gf[kDof], num_lim = T(0);
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        Gd[i] = T(0.2); // base_energy
        Gf[i] = T(0);
#pragma unroll
        for (int up = 0; up < 2; ++up) {
            T x = up ? cfg.lim[i][1] - q[i] : q[i] - cfg.lim[i][0];
            T xd = up ? -qd[i] : qd[i];
            T s = xd > T(0) ? T(0) : (xd < T(0) ? T(1) : T(0.5));
            T ix = Mth<T>::rcp(x);
            T xd2 = xd * xd;
            T Ml = T(0.2) * s * ix;
            T fl = Ml * (T(-0.1) * xd2 * ix);
            T fel = T(-0.1) * s * xd2 * ix * ix;
            Gd[i] += Ml;
            Gf[i] += up ? -fl : fl;
            num_lim += xd * (fl - fel);
        }
    }

    // ---- sphere / plane leaves: lane -> (ego point e, other-robot point pt) ----
    const int e = lane % kEgo, pt = lane / kEgo; // pt < 6 for lanes 0..29
    const bool leaf_lane = lane < kEgo * kPts;
    V3<T> p = kin_load(kin, NT, tid, e, 0), v = kin_load(kin, NT, tid, e, 3), cc = kin_load(kin, NT, tid, e, 6);
    PointAcc<T> acc;
    acc.A = Sym3<T>{T(0), T(0), T(0), T(0), T(0), T(0)};
    acc.b = mk(T(0), T(0), T(0));
    T num = T(0);
    if (cfg.has_coll && leaf_lane) {
        const int rb_first = e + (e > 2 ? 1 : 0);
        T rb = prm[(P_RB + rb_first) * NT + tid];
        T we = T(1);
        int passes = 1;
        if (e == 2) {
            T rb2 = prm[(P_RB + rb_first + 1) * NT + tid];
            if (rb2 == rb) we = T(2); else passes = 2;
        }
        for (int pass = 0; pass < passes; ++pass) {
            if (pass == 1) rb = prm[(P_RB + rb_first + 1) * NT + tid];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (j == r) continue;
                V3<T> xo = kin_load(kin, NT, j, pt, 0), vo = kin_load(kin, NT, j, pt, 3), co = kin_load(kin, NT, j, pt, 6);
                const int nsub = cfg.pt_n[j][pt];
                for (int sub = 0; sub < nsub; ++sub)
                    sphere_leaf(p, v, cc, xo, vo, co, vref, aref, cfg.pt_rad[j][pt][sub] + rb, we * cfg.pt_w[j][pt], sigma,
                                acc, num);
            }
            if (pt == 0) {
                const V3<T> nh = mk(prm[(P_NH + 0) * NT + tid], prm[(P_NH + 1) * NT + tid], prm[(P_NH + 2) * NT + tid]);
                plane_leaf(p, v, cc, nh, prm[P_DN * NT + tid], rb, we, sigma, acc, num);
            }
        }
    }
    // reduce over pt: lanes e, e+5, ..., e+25 -> lane e
    auto red_pt = [&](T x) {
        x += shfl_down_t(x, 15);
        T a = shfl_down_t(x, 5), b = shfl_down_t(x, 10);
        return x + a + b;
    };
    acc.A.xx = red_pt(acc.A.xx); acc.A.xy = red_pt(acc.A.xy); acc.A.xz = red_pt(acc.A.xz);
    acc.A.yy = red_pt(acc.A.yy); acc.A.yz = red_pt(acc.A.yz); acc.A.zz = red_pt(acc.A.zz);
    acc.b.x = red_pt(acc.b.x); acc.b.y = red_pt(acc.b.y); acc.b.z = red_pt(acc.b.z);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) num += shfl_xor_t(num, m); // every lane: sum over all leaves
    num += num_lim;

    // ---- attractor scalars (redundant) and the pull-back operands of each lane ----
    V3<T> org[6];
    org[0] = mk(cfg.link1[r][0], cfg.link1[r][1], cfg.link1[r][2]);
    org[1] = org[0];
    org[2] = kin_load(kin, NT, tid, 0, 0);
    org[3] = kin_load(kin, NT, tid, 1, 0);
    org[4] = kin_load(kin, NT, tid, 2, 0);
    org[5] = org[4];
    V3<T> p8 = kin_load(kin, NT, tid, 4, 0), c8 = kin_load(kin, NT, tid, 4, 6);
    V3<T> p7 = kin_load(kin, NT, tid, 3, 0), c7 = kin_load(kin, NT, tid, 3, 6);
    V3<T> x0 = p8 - mk(prm[(P_G0 + 0) * NT + tid], prm[(P_G0 + 1) * NT + tid], prm[(P_G0 + 2) * NT + tid]);
    T n0 = Mth<T>::sqrt(dot(x0, x0));
    T dpsi0, m0;
    attractor_scalars(n0, prm[P_W0 * NT + tid], dpsi0, m0);
    V3<T> t0 = (x0 * (dpsi0 * Mth<T>::rcp(n0)) + c8 * sigma) * m0;
    T Rg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rg[k] = prm[(P_ANG + k) * NT + tid];
    auto rot = [&](V3<T> u) {
        return V3<T>{Rg[0] * u.x + Rg[1] * u.y + Rg[2] * u.z, Rg[3] * u.x + Rg[4] * u.y + Rg[5] * u.z,
                     Rg[6] * u.x + Rg[7] * u.y + Rg[8] * u.z};
    };
    auto rotT = [&](V3<T> u) {
        return V3<T>{Rg[0] * u.x + Rg[3] * u.y + Rg[6] * u.z, Rg[1] * u.x + Rg[4] * u.y + Rg[7] * u.z,
                     Rg[2] * u.x + Rg[5] * u.y + Rg[8] * u.z};
    };
    V3<T> d87 = p8 - p7;
    V3<T> x1 = rot(d87) - mk(prm[(P_G1 + 0) * NT + tid], prm[(P_G1 + 1) * NT + tid], prm[(P_G1 + 2) * NT + tid]);
    T n1 = Mth<T>::sqrt(dot(x1, x1));
    T dpsi1, m1;
    attractor_scalars(n1, prm[P_W1 * NT + tid], dpsi1, m1);
    V3<T> t1 = (x1 * (dpsi1 * Mth<T>::rcp(n1)) + rot(c8 - c7) * sigma) * m1;

    // lane roles for the pull-back: 0..4 ego points (geometry), 6 sub-goal 0, 7 sub-goal 1, others idle
    int K = 0;
    bool virt = false; // columns z_j x P with P given directly (sub-goal 1: P = hand - link7)
    V3<T> P = p;
    if (lane < kEgo) {
        K = cfg.has_coll ? (lane < 3 ? lane + 2 : 6) : 0;
    } else if (lane == 6) {
        K = 6; P = p8;
        acc.A = Sym3<T>{m0, T(0), T(0), m0, T(0), m0};
        acc.b = t0;
    } else if (lane == 7) {
        K = 6; P = d87; virt = true;
        // J_r^T (m1 I) J_r = Jc^T (m1 R^T R) Jc ;  J_r^T t1 = Jc^T R^T t1
        V3<T> c0 = mk(Rg[0], Rg[3], Rg[6]), c1 = mk(Rg[1], Rg[4], Rg[7]), c2 = mk(Rg[2], Rg[5], Rg[8]);
        acc.A = Sym3<T>{m1 * dot(c0, c0), m1 * dot(c0, c1), m1 * dot(c0, c2), m1 * dot(c1, c1), m1 * dot(c1, c2),
                        m1 * dot(c2, c2)};
        acc.b = rotT(t1);
    }
    T Mp[6][6], fp[6];
    V3<T> Jc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        fp[j] = T(0);
#pragma unroll
        for (int i = 0; i <= j; ++i) Mp[i][j] = T(0);
        if (j < K) Jc[j] = cross(ch.z[j], virt ? P : P - org[j]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        if (j < K) {
            V3<T> AJ = symmul(acc.A, Jc[j]);
            fp[j] = dot(Jc[j], acc.b);
#pragma unroll
            for (int i = 0; i <= j; ++i) Mp[i][j] = dot(Jc[i], AJ);
        }
    }
    // ---- sums: geometry part over lanes 0..4 (8-lane butterfly), attractor part over lanes 6,7 ----
    const bool geo = lane < kEgo, att = (lane == 6 || lane == 7);
    Spec<T> G, F;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        T g = geo ? fp[j] : T(0), a = att ? fp[j] : T(0);
        g += shfl_xor_t(g, 1); g += shfl_xor_t(g, 2); g += shfl_xor_t(g, 4);
        a += shfl_xor_t(a, 1);
        G.f[j] = Gf[j] + shfl_t(g, 0);
        F.f[j] = G.f[j] + shfl_t(a, 6);
#pragma unroll
        for (int i = 0; i <= j; ++i) {
            T gm = geo ? Mp[i][j] : T(0), am = att ? Mp[i][j] : T(0);
            gm += shfl_xor_t(gm, 1); gm += shfl_xor_t(gm, 2); gm += shfl_xor_t(gm, 4);
            am += shfl_xor_t(am, 1);
            G.M[i][j] = shfl_t(gm, 0) + (i == j ? Gd[i] : T(0));
            F.M[i][j] = G.M[i][j] + shfl_t(am, 6);
        }
    }
    G.m7 = Gd[6]; G.f[6] = Gf[6];
    {   // sub-goal 2: joint 7 -> x_goal_2
        T x2 = q[6] - prm[P_G2 * NT + tid];
        T dpsi2, m2;
        attractor_scalars(Mth<T>::abs(x2), prm[P_W2 * NT + tid], dpsi2, m2);
        F.f[6] = G.f[6] + m2 * dpsi2 * (x2 > T(0) ? T(1) : (x2 < T(0) ? T(-1) : T(0)));
        F.m7 = G.m7 + m2;
    }
    T qMq = G.m7 * qd[6] * qd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        qMq += G.M[i][i] * qd[i] * qd[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) qMq += T(2) * G.M[i][j] * qd[i] * qd[j];
    }
    const T a_geom = -num * Mth<T>::rcp(e_ + qMq);
    // ---- even lanes solve the geometry system, odd lanes the forced one (one instruction stream) ----
    const bool odd = lane & 1;
    T Ms[6][6], fs[6], hs[kDof];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        fs[j] = odd ? F.f[j] : G.f[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) Ms[i][j] = odd ? F.M[i][j] : G.M[i][j];
    }
    chol_solve6(Ms, e_, fs, hs);
    hs[6] = (odd ? F.f[6] : G.f[6]) * Mth<T>::rcp((odd ? F.m7 : G.m7) + e_);
    T qq = T(0), qhg = T(0), qhf = T(0), hf[kDof];
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        T hg_i = shfl_t(hs[i], 0);
        hf[i] = shfl_t(hs[i], 1);
        qq += qd[i] * qd[i];
        qhg += qd[i] * hg_i;
        qhf += qd[i] * hf[i];
    }
    const T iden = Mth<T>::rcp(e_ + cfg.s2 * qq);
    const T a_ex0 = -cfg.s2 * qhg * iden, a_exf = -cfg.s2 * qhf * iden;
    const T eta = T(0.5) * (Mth<T>::tanh(T(-0.45) * qq - T(0.5)) + T(1));
    const T a_ex = eta * a_ex0 + (T(1) - eta) * a_exf;
    const T beta = T(0.5) * (Mth<T>::tanh(T(-0.5) * (n0 - T(0.02))) + T(1)) * T(6.5) + T(0.01) +
                   Mth<T>::max(T(0), a_geom - a_ex);
    const T damp = a_ex + beta;
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        T qdd = -hf[i] - damp * qd[i];
        act[i] = cfg.mode == 1 ? qd[i] + cfg.dt * qdd : qdd;
    }
}

template <typename T, int R>
__global__ void __launch_bounds__(32 * R)
    rollout_coop_kernel(const __grid_constant__ DevCfg<T> cfg, const T* __restrict__ rec, int N, T* __restrict__ avg_vel,
                        T* __restrict__ x_ee, T* __restrict__ goal_est, T* __restrict__ qN, T* __restrict__ qdN,
                        long long B) {
    __shared__ T kin[kKin * R];
    __shared__ T prm[P_N * R];
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = blockIdx.x; // one scenario per CTA
    auto ld = [&](int f) { return rec[((long long)f * R + r) * B + b]; };
    T q[kDof], qd[kDof];
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        q[i] = ld(MRF_Q + i);
        qd[i] = ld(MRF_QD + i);
    }
    if (lane == 0) load_params<T>(ld, prm, R, r);
    __syncwarp();
    Chain<T> ch;
    const T vref = cfg.static_or_dyn ? T(1) : T(0), aref = cfg.static_or_dyn ? cfg.sref : T(0);
    T accv = T(0);
    const bool want_pre = x_ee != nullptr || goal_est != nullptr || cfg.estimate_goal != 0;
    for (int k = want_pre ? -1 : 0; k < N; ++k) {
        if (k >= 0) {
#pragma unroll
            for (int i = 0; i < kDof; ++i) q[i] += cfg.dt * qd[i];
        }
        __syncthreads();
        chain_forward(cfg, r, q, qd, ch, kin, R, r); // every lane computes the chain; identical stores
        if (k < 0) {
            __syncwarp();
            V3<T> p8 = kin_load(kin, R, r, 4, 0);
            if (x_ee != nullptr && lane == 0) {
                x_ee[((long long)r * 3 + 0) * B + b] = p8.x;
                x_ee[((long long)r * 3 + 1) * B + b] = p8.y;
                x_ee[((long long)r * 3 + 2) * B + b] = p8.z;
            }
            if (r == cfg.estimate_robot) {
                V3<T> g = mk(prm[(P_G0 + 0) * R + r], prm[(P_G0 + 1) * R + r], prm[(P_G0 + 2) * R + r]);
                if (cfg.estimate_goal) {
                    V3<T> l1 = mk(cfg.link1[r][0], cfg.link1[r][1], cfg.link1[r][2]);
                    V3<T> v = cfg.estimate_goal == 1 ? cross(ch.z[0], p8 - l1) : kin_load(kin, R, r, 4, 3);
                    g = p8 + v * cfg.est_h;
                    __syncwarp();
                    if (lane == 0) {
                        prm[(P_G0 + 0) * R + r] = g.x;
                        prm[(P_G0 + 1) * R + r] = g.y;
                        prm[(P_G0 + 2) * R + r] = g.z;
                    }
                    __syncwarp();
                }
                if (goal_est != nullptr && lane == 0) {
                    goal_est[0 * B + b] = g.x;
                    goal_est[1 * B + b] = g.y;
                    goal_est[2 * B + b] = g.z;
                }
            }
            continue;
        }
        __syncthreads();
        T act[kDof];
        fabric_action_coop<T, R>(cfg, r, lane, q, qd, ch, kin, prm, vref, aref, act);
#pragma unroll
        for (int i = 0; i < kDof; ++i) {
            qd[i] = act[i];
            accv += act[i] * act[i];
        }
        if (lane == 0) {
            if (qN != nullptr) {
#pragma unroll
                for (int i = 0; i < kDof; ++i) qN[(((long long)r * N + k) * kDof + i) * B + b] = q[i];
            }
            if (qdN != nullptr) {
#pragma unroll
                for (int i = 0; i < kDof; ++i) qdN[(((long long)r * N + k) * kDof + i) * B + b] = qd[i];
            }
        }
    }
    if (avg_vel != nullptr && lane == 0) avg_vel[(long long)r * B + b] = accv / (T(N) * T(kDof));
}

} // namespace mrf
