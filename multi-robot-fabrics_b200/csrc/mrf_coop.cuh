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

// Per-robot scratch in shared memory (besides the published points and the parameter block)
template <typename T> struct CoopState {
    T q[kDof], qd[kDof]; // authoritative joint state of the robot
    T sn[kDof], cs[kDof]; // sin / cos of the joint angles (computed by lanes 0..6)
    T z[6][3];           // world axes of joints 1..6
};

// Panda chain as a loop (one copy of the per-joint code: the cooperative kernel is instruction-fetch bound, its step
// body has to stay inside the 32 KB instruction cache).  Segment i: advance by offset i (in the frame before joint
// i's origin roll), publish the point if it is a collision point, then apply joint i.
__constant__ float kChainOff[8][3] = {{0.f, 0.f, 0.f},         {0.f, 0.f, 0.f},  {0.f, -0.316f, 0.f}, {0.0825f, 0.f, 0.f},
                                      {-0.0825f, 0.384f, 0.f}, {0.f, 0.f, 0.f},  {0.088f, 0.f, 0.f},  {0.f, 0.f, 0.107f}};
__constant__ int kChainSlot[8] = {-1, -1, 0, 1, 2, -1, 3, 4};
__constant__ int kChainRoll[8] = {0, -1, 1, 1, -1, 1, 1, 0};
__constant__ double kChainOffD[8][3] = {{0., 0., 0.},         {0., 0., 0.},  {0., -0.316, 0.}, {0.0825, 0., 0.},
                                        {-0.0825, 0.384, 0.}, {0., 0., 0.},  {0.088, 0., 0.},  {0., 0., 0.107}};
template <typename T> __device__ __forceinline__ T chain_off(int i, int c);
template <> __device__ __forceinline__ float chain_off<float>(int i, int c) { return kChainOff[i][c]; }
template <> __device__ __forceinline__ double chain_off<double>(int i, int c) { return kChainOffD[i][c]; }

template <typename T, int R>
__device__ __forceinline__ void chain_forward_loop(const DevCfg<T>& cfg, int r, CoopState<T>& st, T* kin) {
    Frame<T> f;
    const T* Rm = cfg.R0[r];
    f.a = mk(Rm[0], Rm[3], Rm[6]);
    f.b = mk(Rm[1], Rm[4], Rm[7]);
    f.n = mk(Rm[2], Rm[5], Rm[8]);
    f.p = mk(cfg.link1[r][0], cfg.link1[r][1], cfg.link1[r][2]);
    f.w = mk(T(0), T(0), T(0));
    f.al = f.w; f.v = f.w; f.ac = f.w;
    kin_store(kin, R, r, 5, f); // link1 == link2
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int slot = kChainSlot[i];
        if (slot >= 0) {
            fr_advance(f, f.a * chain_off<T>(i, 0) + f.b * chain_off<T>(i, 1) + f.n * chain_off<T>(i, 2));
            kin_store(kin, R, r, slot, f);
        }
        if (i < 7) {
            const int roll = kChainRoll[i];
            if (roll != 0) {
                const T rs = T(roll);
                V3<T> t = f.b;
                f.b = f.n * rs;
                f.n = t * (-rs);
            }
            const T s = st.sn[i], c = st.cs[i];
            V3<T> a = f.a * c + f.b * s;
            V3<T> b = f.b * c - f.a * s;
            f.a = a;
            f.b = b;
            V3<T> zq = f.n * st.qd[i];
            f.al = f.al + cross(f.w, zq);
            f.w = f.w + zq;
            if (i < 6) {
                st.z[i][0] = f.n.x; st.z[i][1] = f.n.y; st.z[i][2] = f.n.z;
            }
        }
    }
}

// Phase B of one robot, executed by its warp.  kin / prm are indexed [k * R + r] (one column per robot).
// Lane roles: leaves  lane = e + 5 pt (e ego point, pt other-robot point), limits lane < 14;
//             pull-back group G = lanes 0..7 (0..4 ego points), group F = lanes 8..15 (8..12 ego points again,
//             13 sub-goal 0, 14 sub-goal 1); a shuffle-down tree leaves M_G, f_G on lane 0 and M_F, f_F on lane 8.
template <typename T, int R>
__device__ __forceinline__ void fabric_action_coop(const DevCfg<T>& cfg, int r, int lane, const CoopState<T>& st,
                                                   const T* kin, const T* prm, T vref, T aref, T* act) {
    constexpr int NT = R;
    const int tid = r;
    const T sigma = cfg.sigma;
    const T e_ = cfg.eps;
    T num = T(0);
    // ---- joint-limit leaves: lane -> (joint, side) ----
    T lim_M = T(0), lim_f = T(0);
    if (lane < 2 * kDof) {
        const int i = lane >> 1, up = lane & 1;
        T x = up ? cfg.lim[i][1] - st.q[i] : st.q[i] - cfg.lim[i][0];
        T xd = up ? -st.qd[i] : st.qd[i];
        T s = xd > T(0) ? T(0) : (xd < T(0) ? T(1) : T(0.5));
        T ix = Mth<T>::rcp(x);
        T xd2 = xd * xd;
        lim_M = T(0.2) * s * ix;
        T fl = lim_M * (T(-0.1) * xd2 * ix);
        T fel = T(-0.1) * s * xd2 * ix * ix;
        lim_f = up ? -fl : fl;
        num = xd * (fl - fel);
    }
    lim_M += shfl_xor_t(lim_M, 1); // lanes 2i, 2i+1: joint i
    lim_f += shfl_xor_t(lim_f, 1);

    // ---- sphere / plane leaves: lane -> (ego point e, other-robot point pt) ----
    const int e = lane % kEgo, pt = lane / kEgo; // pt < 6 for lanes 0..29
    V3<T> p = kin_load(kin, NT, tid, e, 0), v = kin_load(kin, NT, tid, e, 3), cc = kin_load(kin, NT, tid, e, 6);
    PointAcc<T> acc;
    acc.A = Sym3<T>{T(0), T(0), T(0), T(0), T(0), T(0)};
    acc.b = mk(T(0), T(0), T(0));
    if (cfg.has_coll && lane < kEgo * kPts) {
        const int rb_first = e + (e > 2 ? 1 : 0);
        T rb = prm[(P_RB + rb_first) * NT + tid];
        T we = T(1);
        int passes = 1;
        const int em = cfg.ego_mask[r]; // which of panda_link3..8 are collision links (see fabric_action)
        if (e == 2) {
            T rb2 = prm[(P_RB + rb_first + 1) * NT + tid];
            if (((em >> rb_first) & 3) == 3) {
                if (rb2 == rb) we = T(2); else passes = 2;
            } else if ((em >> (rb_first + 1)) & 1) {
                rb = rb2;
            } else if (!((em >> rb_first) & 1)) {
                passes = 0;
            }
        } else if (!((em >> rb_first) & 1)) {
            passes = 0;
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
    // reduce over pt (lanes e, e+5, ..., e+25 -> lane e), then hand lanes 8..12 a copy for the F group
    auto red_pt = [&](T x) {
        x += shfl_down_t(x, 15);
        T a = shfl_down_t(x, 5), b = shfl_down_t(x, 10);
        return shfl_t(x + a + b, lane & 7);
    };
    acc.A.xx = red_pt(acc.A.xx); acc.A.xy = red_pt(acc.A.xy); acc.A.xz = red_pt(acc.A.xz);
    acc.A.yy = red_pt(acc.A.yy); acc.A.yz = red_pt(acc.A.yz); acc.A.zz = red_pt(acc.A.zz);
    acc.b.x = red_pt(acc.b.x); acc.b.y = red_pt(acc.b.y); acc.b.z = red_pt(acc.b.z);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) num += shfl_xor_t(num, m); // every lane: sum over all leaves

    // ---- pull-back operands.  Lanes 13 / 14 build the two task-space attractors with one code path:
    //      x = Rm (P) - goal, A = m Rm^T Rm, b = Rm^T m (dpsi x/|x| + sigma Rm c)  with Rm = I or angle_goal_1 ----
    V3<T> org[6];
    org[0] = mk(cfg.link1[r][0], cfg.link1[r][1], cfg.link1[r][2]);
    org[1] = org[0];
    org[2] = kin_load(kin, NT, tid, 0, 0);
    org[3] = kin_load(kin, NT, tid, 1, 0);
    org[4] = kin_load(kin, NT, tid, 2, 0);
    org[5] = org[4];
    const int gl = lane & 7;           // role inside the 8-lane group
    const bool second = (lane & 8) != 0; // F group
    int K = 0;
    bool virt = false;
    V3<T> P = kin_load(kin, NT, tid, gl < kEgo ? gl : 0, 0);
    T xnorm = T(0);
    if (gl < kEgo) {
        K = cfg.has_coll ? (gl < 3 ? gl + 2 : 6) : 0;
    }
    if (second && (gl == 5 || gl == 6)) {
        const bool g1 = gl == 6;
        V3<T> p8 = kin_load(kin, NT, tid, 4, 0), c8 = kin_load(kin, NT, tid, 4, 6);
        T Rg[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rg[k] = g1 ? prm[(P_ANG + k) * NT + tid] : ((k == 0 || k == 4 || k == 8) ? T(1) : T(0));
        V3<T> cterm = c8;
        P = p8;
        if (g1) {
            V3<T> p7 = kin_load(kin, NT, tid, 3, 0), c7 = kin_load(kin, NT, tid, 3, 6);
            P = p8 - p7;
            cterm = c8 - c7;
        }
        const int go = g1 ? P_G1 : P_G0;
        auto rot = [&](V3<T> u) {
            return V3<T>{Rg[0] * u.x + Rg[1] * u.y + Rg[2] * u.z, Rg[3] * u.x + Rg[4] * u.y + Rg[5] * u.z,
                         Rg[6] * u.x + Rg[7] * u.y + Rg[8] * u.z};
        };
        V3<T> x = rot(P) - mk(prm[(go + 0) * NT + tid], prm[(go + 1) * NT + tid], prm[(go + 2) * NT + tid]);
        xnorm = Mth<T>::sqrt(dot(x, x));
        T dpsi, m2;
        attractor_scalars(xnorm, prm[(g1 ? P_W1 : P_W0) * NT + tid], dpsi, m2);
        V3<T> t = (x * (dpsi * Mth<T>::rcp(xnorm)) + rot(cterm) * sigma) * m2;
        V3<T> c0 = mk(Rg[0], Rg[3], Rg[6]), c1 = mk(Rg[1], Rg[4], Rg[7]), c2 = mk(Rg[2], Rg[5], Rg[8]);
        acc.A = Sym3<T>{m2 * dot(c0, c0), m2 * dot(c0, c1), m2 * dot(c0, c2), m2 * dot(c1, c1), m2 * dot(c1, c2),
                        m2 * dot(c2, c2)};
        acc.b = mk(dot(c0, t), dot(c1, t), dot(c2, t));
        K = 6;
        virt = g1; // sub-goal 1: columns z_j x (hand - link7)
    }
    const T xpsi = shfl_t(xnorm, 13); // |hand - x_goal_0| for the damper
    T Mp[6][6], fp[6];
    V3<T> Jc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        fp[j] = T(0);
#pragma unroll
        for (int i = 0; i <= j; ++i) Mp[i][j] = T(0);
        if (j < K) Jc[j] = cross(mk(st.z[j][0], st.z[j][1], st.z[j][2]), virt ? P : P - org[j]);
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
    // ---- 8-lane shuffle-down trees: lane 0 = geometry sums, lane 8 = geometry + attractor sums ----
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        fp[j] += shfl_down_t(fp[j], 4); fp[j] += shfl_down_t(fp[j], 2); fp[j] += shfl_down_t(fp[j], 1);
#pragma unroll
        for (int i = 0; i <= j; ++i) {
            Mp[i][j] += shfl_down_t(Mp[i][j], 4); Mp[i][j] += shfl_down_t(Mp[i][j], 2); Mp[i][j] += shfl_down_t(Mp[i][j], 1);
        }
    }
    // base energy + limit leaves on the diagonal (both systems), sub-goal 2 on joint 7 of the forced system
    T m7 = T(0.2) + shfl_t(lim_M, 12), f7 = shfl_t(lim_f, 12);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        Mp[i][i] += T(0.2) + shfl_t(lim_M, 2 * i);
        fp[i] += shfl_t(lim_f, 2 * i);
    }
    const T m7g = m7, f7g = f7;
    {
        T x2 = st.q[6] - prm[P_G2 * NT + tid];
        T dpsi2, m2;
        attractor_scalars(Mth<T>::abs(x2), prm[P_W2 * NT + tid], dpsi2, m2);
        if (second) {
            f7 += m2 * dpsi2 * (x2 > T(0) ? T(1) : (x2 < T(0) ? T(-1) : T(0)));
            m7 += m2;
        }
    }
    T qd[kDof];
#pragma unroll
    for (int i = 0; i < kDof; ++i) qd[i] = st.qd[i];
    T qMq = m7g * qd[6] * qd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        qMq += Mp[i][i] * qd[i] * qd[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) qMq += T(2) * Mp[i][j] * qd[i] * qd[j];
    }
    qMq = shfl_t(qMq, 0); // geometry metric lives on lane 0
    (void)f7g;
    const T a_geom = -num * Mth<T>::rcp(e_ + qMq);
    // ---- lane 0 solves the geometry system, lane 8 the forced one (same instruction stream) ----
    T hs[kDof];
    chol_solve6(Mp, e_, fp, hs);
    hs[6] = f7 * Mth<T>::rcp(m7 + e_);
    T qq = T(0), qhg = T(0), qhf = T(0), hf[kDof];
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        T hg_i = shfl_t(hs[i], 0);
        hf[i] = shfl_t(hs[i], 8);
        qq += qd[i] * qd[i];
        qhg += qd[i] * hg_i;
        qhf += qd[i] * hf[i];
    }
    const T iden = Mth<T>::rcp(e_ + cfg.s2 * qq);
    const T a_ex0 = -cfg.s2 * qhg * iden, a_exf = -cfg.s2 * qhf * iden;
    const T eta = T(0.5) * (Mth<T>::tanh(T(-0.45) * qq - T(0.5)) + T(1));
    const T a_ex = eta * a_ex0 + (T(1) - eta) * a_exf;
    const T beta = T(0.5) * (Mth<T>::tanh(T(-0.5) * (xpsi - T(0.02))) + T(1)) * T(6.5) + T(0.01) +
                   Mth<T>::max(T(0), a_geom - a_ex);
    const T damp = a_ex + beta;
#pragma unroll
    for (int i = 0; i < kDof; ++i) {
        T qdd = -hf[i] - damp * qd[i];
        act[i] = cfg.mode == 1 ? qd[i] + cfg.dt * qdd : qdd;
    }
}

template <typename T, int R>
__global__ void __launch_bounds__(32 * R, 1)
    rollout_coop_kernel(const __grid_constant__ DevCfg<T> cfg, const T* __restrict__ rec, int N, T* __restrict__ avg_vel,
                        T* __restrict__ x_ee, T* __restrict__ goal_est, T* __restrict__ qN, T* __restrict__ qdN,
                        long long B) {
    __shared__ T kin[kKinRows<T> * R];
    __shared__ T prm[P_N * R];
    __shared__ CoopState<T> sts[R];
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = blockIdx.x; // one scenario per CTA
    auto ld = [&](int f) { return rec[((long long)f * R + r) * B + b]; };
    CoopState<T>& st = sts[r];
    if (lane < kDof) {
        st.q[lane] = ld(MRF_Q + lane);
        st.qd[lane] = ld(MRF_QD + lane);
    }
    if (lane == 31) load_params<T>(ld, prm, R, r);
    __syncwarp();
    const T vref = cfg.static_or_dyn ? T(1) : T(0), aref = cfg.static_or_dyn ? cfg.sref : T(0);
    T accv = T(0);
    const bool want_pre = x_ee != nullptr || goal_est != nullptr || cfg.estimate_goal != 0;
#pragma unroll 1
    for (int k = want_pre ? -1 : 0; k < N; ++k) {
        if (lane < kDof) { // lanes 0..6 own one joint each: integrate and take sin / cos
            T qn = st.q[lane];
            if (k >= 0) qn += cfg.dt * st.qd[lane];
            T s, c;
            Mth<T>::sincos(qn, &s, &c);
            st.q[lane] = qn;
            st.sn[lane] = s;
            st.cs[lane] = c;
        }
        __syncthreads(); // state visible to the warp; readers of the previous step's points are done
        chain_forward_loop<T, R>(cfg, r, st, kin); // every lane runs the chain; identical stores
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
                    V3<T> z0 = mk(st.z[0][0], st.z[0][1], st.z[0][2]);
                    V3<T> v = cfg.estimate_goal == 1 ? cross(z0, p8 - l1) : kin_load(kin, R, r, 4, 3);
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
        __syncthreads(); // every robot of the scenario has published
        T act[kDof];
        fabric_action_coop<T, R>(cfg, r, lane, st, kin, prm, vref, aref, act);
        __syncwarp(); // all lanes have read the old velocities
#pragma unroll
        for (int i = 0; i < kDof; ++i) accv += act[i] * act[i];
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kDof; ++i) st.qd[i] = act[i];
            if (qN != nullptr) {
#pragma unroll
                for (int i = 0; i < kDof; ++i) qN[(((long long)r * N + k) * kDof + i) * B + b] = st.q[i];
            }
            if (qdN != nullptr) {
#pragma unroll
                for (int i = 0; i < kDof; ++i) qdN[(((long long)r * N + k) * kDof + i) * B + b] = act[i];
            }
        }
        __syncwarp();
    }
    if (avg_vel != nullptr && lane == 0) avg_vel[(long long)r * B + b] = accv / (T(N) * T(kDof));
}

} // namespace mrf
