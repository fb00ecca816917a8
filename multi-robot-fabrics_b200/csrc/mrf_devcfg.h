// mrf_devcfg.h -- host-side translation of the public MrfConfig into the kernels' DevCfg<T>.
#pragma once
#include <cstring>

#include "../../include/mrf_b200.h"
#include "mrf_device.cuh"

namespace mrf {
template <typename T> inline void fill_devcfg(const MrfConfig& c, DevCfg<T>& d) {
    memset(&d, 0, sizeof(d));
    d.n_robots = c.n_robots;
    d.mode = c.mode;
    d.static_or_dyn = c.static_or_dyn;
    d.has_coll = c.has_collision_links;
    d.estimate_goal = c.estimate_goal;
    d.estimate_robot = (c.estimate_robot >= 0 && c.estimate_robot < c.n_robots) ? c.estimate_robot : -1;
    d.est_h = (T)c.estimate_horizon;
    d.dt = (T)c.dt;
    d.eps = (T)c.eps;
    d.sigma = (T)c.jdot_sign;
    d.sref = (T)c.jdot_ref_sign;
    d.s2 = (T)(2.0 * c.exec_scale);
    for (int r = 0; r < MRF_MAX_ROBOTS; ++r) {
        const double* M = c.mount[r];
        const double R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
        for (int k = 0; k < 9; ++k) d.R0[r][k] = (T)R[k];
        d.p0[r][0] = (T)M[3];
        d.p0[r][1] = (T)M[7];
        d.p0[r][2] = (T)M[11];
        // panda_joint1 origin (0,0,0.333), panda_with_finger.urdf:99
        d.link1[r][0] = (T)(M[3] + R[2] * 0.333);
        d.link1[r][1] = (T)(M[7] + R[5] * 0.333);
        d.link1[r][2] = (T)(M[11] + R[8] * 0.333);
        for (int l = 0; l < MRF_NLINKS; ++l) d.r_link[r][l] = (T)c.r_robots[r][l];
    }
    // collision_link_mask: bit l-1 <=> panda_link l is a collision link of robot r (collision_links_nr of
    // set_planner_panda).  As EGO links only link3..8 count (constant fk below); as obstacles of the others all do.
    auto on = [&](int r, int l) { return ((c.collision_link_mask[r] >> l) & 1) != 0; };
    for (int r = 0; r < MRF_MAX_ROBOTS; ++r) d.ego_mask[r] = c.has_collision_links ? (c.collision_link_mask[r] >> 2) & 0x3F : 0;
    for (int r = 0; r < c.n_robots; ++r) {
        const int la[kPts] = {2, 3, 4, 6, 7, 0}, lb[kPts] = {-1, -1, 5, -1, -1, 1}; // link indices sharing each point
        for (int pt = 0; pt < kPts; ++pt) {
            const bool a_on = on(r, la[pt]), b_on = lb[pt] >= 0 && on(r, lb[pt]);
            const bool same = a_on && b_on && c.r_robots[r][la[pt]] == c.r_robots[r][lb[pt]];
            d.pt_rad[r][pt][0] = (T)c.r_robots[r][a_on ? la[pt] : (b_on ? lb[pt] : la[pt])];
            d.pt_rad[r][pt][1] = (T)(b_on ? c.r_robots[r][lb[pt]] : 0.0);
            d.pt_n[r][pt] = (a_on && b_on && !same) ? 2 : ((a_on || b_on) ? 1 : 0);
            d.pt_w[r][pt] = (T)(same ? 2 : 1);
        }
    }
    d.uniform_obst = 1;
    d.r_obst = (T)c.r_robots[0][0];
    for (int r = 0; r < c.n_robots; ++r) {
        if ((c.collision_link_mask[r] & 0xFF) != 0xFF) d.uniform_obst = 0;
        for (int l = 0; l < 8; ++l)
            if (c.r_robots[r][l] != c.r_robots[0][0]) d.uniform_obst = 0;
    }
    // sphere table of robot r: links of every other robot j (ascending), (1,2) -> point 5, 3 -> 0, 4 -> 1,
    // (5,6) -> 2, 7 -> 3, 8 -> 4
    for (int r = 0; r < c.n_robots; ++r) {
        const int src_of[8] = {5, 5, 0, 1, 2, 2, 3, 4};
        int n = 0;
        for (int j = 0; j < c.n_robots; ++j) {
            if (j == r) continue;
            for (int l = 0; l < 8; ++l) {
                if (!on(j, l)) continue;
                if ((l == 1 || l == 5) && on(j, l - 1) && c.r_robots[j][l] == c.r_robots[j][l - 1]) {
                    d.ent_w[r][n - 1] = (T)2; // same point and radius as the previous link: one entry, weight 2
                    continue;
                }
                d.ent_rob[r][n] = j;
                d.ent_src[r][n] = src_of[l];
                d.ent_rad[r][n] = (T)c.r_robots[j][l];
                d.ent_w[r][n] = (T)1;
                ++n;
            }
        }
        d.ent_n[r] = n;
    }
    for (int i = 0; i < MRF_DOF; ++i) {
        d.lim[i][0] = (T)c.limits[i][0];
        d.lim[i][1] = (T)c.limits[i][1];
    }
}


} // namespace mrf
