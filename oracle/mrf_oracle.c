/* Oracle O2 -- closed-form float64 CPU restatement.  See mrf_oracle.h for scope and the
 * "parity unpinned" statement.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows, in the reference (/root/reference):
 *   planner configuration   examples/example_pandas_Jointspace.py:25-62 (goal), :84-90 (strings),
 *                           :97-105 (limits), :108-118 (mount), :123-133 (components, 'vel', dt 0.01)
 *   kinematics helpers      multi_robot_fabrics/utils/utils.py:16-54 (fk, J, Jdot with Jdot_sign=-1)
 *   coupled rollout         multi_robot_fabrics/fabrics_planner/forward_planner_Jointspace.py:72-116,190-249
 *   decoupled rollout       multi_robot_fabrics/fabrics_planner/forward_planner_Cartesian.py:77-92,421-458
 *   Panda chain constants   examples/simulation_environments/urdfs/panda_with_finger.urdf:98-470
 * and the fabrics 0.9.5 algorithm (third-party, restated):
 *   leaf spec  M = d2L/dxdot2, f = M h;  pull  M_q = J^T M J, f_q = J^T (f + M Jdot qdot) with
 *   Jdot = jdot_sign * d(J qdot)/dq;  dynamic pull f -= M xddot_ref;  energies pulled by substitution
 *   (exact Euler-Lagrange in q);  qdd = -(M_f+eps I)^-1 f_f - (a_ex + beta) qdot.
 *
 * Written as plain explicit-Jacobian loops (every leaf pulled straight into joint space); the CUDA
 * kernels use a different factorisation (task-space accumulation per link), so agreement between
 * the two is a meaningful check.  Compile: see oracle/Makefile (-O2 -ffp-contract=off).
 */
#include "mrf_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DOF MRFO_DOF

/* ---- Panda chain: joint origins (xyz) and the roll of each origin (all rpy = (r,0,0)) ---- */
static const double JXYZ[7][3] = {{0, 0, 0.333},      {0, 0, 0},     {0, -0.316, 0}, {0.0825, 0, 0},
                                  {-0.0825, 0.384, 0}, {0, 0, 0},     {0.088, 0, 0}};
static const double JROLL[7] = {0.0, -M_PI_2, M_PI_2, M_PI_2, -M_PI_2, M_PI_2, M_PI_2};
static const double LINK8_Z = 0.107; /* fixed panda_joint8; panda_hand origin == panda_link8 origin */

static void cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matmul3(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
            C[3 * i + j] = s;
        }
}

void mrfo_config_default(mrfo_config* c, int n_robots) {
    memset(c, 0, sizeof(*c));
    c->n_robots = n_robots;
    c->mode = 1;
    c->static_or_dyn = 1;
    c->has_collision_links = 1;
    c->dt = 0.01;
    c->eps = 1e-6;
    c->jdot_sign = -1.0;
    c->jdot_ref_sign = -1.0;
    c->exec_scale = 1.0;
    /* parameters_manipulators.py:83-110,138-150: mounts (0,0,.65) yaw 0; (1,0,.65) yaw pi; (0.7,0.6,.65) yaw pi */
    static const double pos[3][3] = {{0.0, 0.0, 0.65}, {1.0, 0.0, 0.65}, {0.7, 0.6, 0.65}};
    for (int r = 0; r < MRFO_MAX_ROBOTS; r++) {
        double yaw = (r == 1 || r == 2) ? M_PI : 0.0; /* set_planner_panda: i_robot in {1, 2} */
        const double* p = pos[r < 3 ? r : 0];
        double* T = c->mount[r];
        memset(T, 0, 16 * sizeof(double));
        T[0] = cos(yaw); T[1] = -sin(yaw); T[4] = sin(yaw); T[5] = cos(yaw); T[10] = 1; T[15] = 1;
        T[3] = p[0]; T[7] = p[1]; T[11] = p[2];
        for (int l = 0; l < MRFO_NLINKS; l++) c->r_robots[r][l] = 0.08; /* parameters_manipulators.py:23 */
        c->link_mask[r] = 0xFF;                                            /* collision_links_nrs = [1..8], :25 */
    }
    static const double lim[7][2] = {{-2.8973, 2.8973}, {-1.7628, 1.7628}, {-2.8973, 2.8973}, {-3.0718, -0.0698},
                                     {-2.8973, 2.8973}, {-0.0175, 3.7525}, {-2.8973, 2.8973}};
    memcpy(c->limits, lim, sizeof(lim));
}

/* ------------------------------------------------------------------------------------------- */
/* Kinematics.  Frame i = panda_link(i+1).  p[i] origin, z[i] joint axis (world), explicit J.  */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    double Rl[8][9];   /* rotation of the link frames link1..8 (after the joint rotation; link8 = link7) */
    double wl[8][3];   /* angular velocity of the links */
    double p[8][3];    /* link origins link1..8 */
    double z[7][3];    /* joint axes */
    double R7[9];      /* rotation of link7 frame */
    double v[8][3];    /* J qdot */
    double c[8][3];    /* TRUE  d(J qdot)/dq qdot */
    double J[8][3][7]; /* positional Jacobians */
} kin_t;

static void kinematics(const double T0[16], const double* q, const double* qd, kin_t* k) {
    double R[9] = {T0[0], T0[1], T0[2], T0[4], T0[5], T0[6], T0[8], T0[9], T0[10]};
    double p[3] = {T0[3], T0[7], T0[11]};
    double w[3] = {0, 0, 0}, vl[3] = {0, 0, 0}, al[3] = {0, 0, 0}, ac[3] = {0, 0, 0}; /* omega, v, alpha, a */
    for (int i = 0; i < 7; i++) {
        /* p_i = p_{i-1} + R_{i-1} xyz_i ; point rigidly attached to frame i-1 */
        double r[3], t[3], u[3];
        for (int a = 0; a < 3; a++) r[a] = R[3 * a] * JXYZ[i][0] + R[3 * a + 1] * JXYZ[i][1] + R[3 * a + 2] * JXYZ[i][2];
        cross(w, r, t);
        for (int a = 0; a < 3; a++) vl[a] += t[a];
        cross(al, r, u);
        double wt[3];
        cross(w, t, wt);
        for (int a = 0; a < 3; a++) ac[a] += u[a] + wt[a];
        for (int a = 0; a < 3; a++) p[a] += r[a];
        /* R_i = R_{i-1} Rx(roll_i) Rz(q_i) */
        double cr = cos(JROLL[i]), sr = sin(JROLL[i]), cq = cos(q[i]), sq = sin(q[i]);
        double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr}, Rz[9] = {cq, -sq, 0, sq, cq, 0, 0, 0, 1}, A[9];
        matmul3(R, Rx, A);
        matmul3(A, Rz, R);
        for (int a = 0; a < 3; a++) {
            k->p[i][a] = p[a];
            k->z[i][a] = R[3 * a + 2];
            k->v[i][a] = vl[a];
            k->c[i][a] = ac[a];
        }
        /* omega_i = omega_{i-1} + z_i qd_i ; alpha_i = alpha_{i-1} + omega_{i-1} x z_i qd_i  (qdd = 0) */
        double zq[3] = {k->z[i][0] * qd[i], k->z[i][1] * qd[i], k->z[i][2] * qd[i]}, wz[3];
        cross(w, zq, wz);
        for (int a = 0; a < 3; a++) {
            al[a] += wz[a];
            w[a] += zq[a];
        }
        memcpy(k->Rl[i], R, sizeof(R));
        memcpy(k->wl[i], w, sizeof(w));
    }
    memcpy(k->Rl[7], R, sizeof(R));
    memcpy(k->wl[7], w, sizeof(w));
    memcpy(k->R7, R, sizeof(R));
    { /* link8 = link7 + R7 (0,0,0.107) */
        double r[3] = {R[2] * LINK8_Z, R[5] * LINK8_Z, R[8] * LINK8_Z}, t[3], u[3], wt[3];
        cross(w, r, t);
        cross(al, r, u);
        cross(w, t, wt);
        for (int a = 0; a < 3; a++) {
            k->p[7][a] = p[a] + r[a];
            k->v[7][a] = vl[a] + t[a];
            k->c[7][a] = ac[a] + u[a] + wt[a];
        }
    }
    for (int l = 0; l < 8; l++)
        for (int j = 0; j < 7; j++) {
            double col[3] = {0, 0, 0};
            if (j <= l) { /* joint j+1 moves link l+1 iff j <= l (link8 rides on joint 7) */
                double d[3] = {k->p[l][0] - k->p[j][0], k->p[l][1] - k->p[j][1], k->p[l][2] - k->p[j][2]};
                cross(k->z[j], d, col);
            }
            for (int a = 0; a < 3; a++) k->J[l][a][j] = col[a];
        }
}

void mrfo_kinematics(const mrfo_config* c, int robot, const double* q, const double* qd, double x[8][3],
                     double v[8][3], double cdd[8][3], double J[8][3][7]) {
    kin_t k;
    kinematics(c->mount[robot], q, qd, &k);
    memcpy(x, k.p, sizeof(k.p));
    memcpy(v, k.v, sizeof(k.v));
    memcpy(cdd, k.c, sizeof(k.c));
    if (J) memcpy(J, k.J, sizeof(k.J));
}

void mrfo_endeffector(const mrfo_config* c, int robot, const double* q, const double* qd, int use_jqd,
                      double x_ee[3], double v_ee[3]) {
    kin_t k;
    kinematics(c->mount[robot], q, qd, &k);
    for (int a = 0; a < 3; a++) {
        x_ee[a] = k.p[7][a];
        v_ee[a] = use_jqd ? k.v[7][a] : k.J[7][a][0];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* Leaf accumulation in joint space                                                            */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    double M[DOF][DOF];
    double f[DOF];
    double fe[DOF];
} spec_t;

/* 1-D leaf with root Jacobian row j, leaf metric Ml, leaf force fl (= Ml h), leaf energy force fel,
 * true curvature `curv` (d2x/dt2 at qdd = 0 without reference acceleration) and reference term acc_o. */
static void add_leaf(spec_t* s, const double j[DOF], double Ml, double fl, double fel, double curv, double acc_o,
                     double sigma) {
    double fq = fl + Ml * (sigma * curv - acc_o);
    double feq = fel + Ml * (curv - acc_o);
    for (int a = 0; a < DOF; a++) {
        s->f[a] += j[a] * fq;
        s->fe[a] += j[a] * feq;
        for (int b = 0; b < DOF; b++) s->M[a][b] += Ml * j[a] * j[b];
    }
}

static double sw(double xd) { return -0.5 * ((xd > 0) - (xd < 0) - 1.0); } /* -0.5 (sign(xd) - 1) */

static int cholesky_solve(const double A[DOF][DOF], double eps, const double b[DOF], double x[DOF]) {
    double L[DOF][DOF];
    memset(L, 0, sizeof(L));
    for (int i = 0; i < DOF; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j] + (i == j ? eps : 0.0);
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0)) return 1;
                L[i][i] = sqrt(s);
            } else
                L[i][j] = s / L[j][j];
        }
    double y[DOF];
    for (int i = 0; i < DOF; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
        y[i] = s / L[i][i];
    }
    for (int i = DOF - 1; i >= 0; i--) {
        double s = y[i];
        for (int k = i + 1; k < DOF; k++) s -= L[k][i] * x[k];
        x[i] = s / L[i][i];
    }
    return 0;
}

static int action_from_kin(const mrfo_config* c, int link_mask, const kin_t* k, const double* rec, int S, const double* xo,
                           const double* vo, const double* ao, const double* ro, double* action, double* diag) {
    const double* q = rec + MRFO_Q;
    const double* qd = rec + MRFO_QD;
    const double sigma = c->jdot_sign;
    spec_t g;
    memset(&g, 0, sizeof(g));
    for (int a = 0; a < DOF; a++) g.M[a][a] = 0.2; /* base_energy 0.5*0.2*qd.qd : M = 0.2 I, h = 0, f_e = 0 */

    if (c->has_collision_links) {
        const double* con = rec + MRFO_CON;
        double nn = sqrt(con[0] * con[0] + con[1] * con[1] + con[2] * con[2]);
        double nh[3] = {con[0] / nn, con[1] / nn, con[2] / nn};
        for (int l = 2; l < 8; l++) { /* panda_link3..8: links 1,2 have constant fk and are skipped */
            if (!((link_mask >> l) & 1)) continue; /* not in collision_links_nr (example_pandas_Jointspace.py:91-96) */
            double rb = rec[MRFO_RB + (l - 2)];
            const double* p = k->p[l];
            const double* v = k->v[l];
            const double* cl = k->c[l];
            for (int o = 0; o < S; o++) { /* dynamic sphere leaves */
                double d[3], w[3], gh[3], j[DOF];
                for (int a = 0; a < 3; a++) {
                    d[a] = p[a] - xo[3 * o + a];
                    w[a] = v[a] - vo[3 * o + a];
                }
                double n = sqrt(dot3(d, d)), rho = ro[o] + rb;
                double x = n / rho - 1.0;
                for (int a = 0; a < 3; a++) gh[a] = d[a] / (n * rho);
                double xd = dot3(gh, w);
                double dw = dot3(d, w);
                double kappa = (dot3(w, w) - dw * dw / (n * n)) / (n * rho);
                double x2 = x * x, x4 = x2 * x2;
                double Ml = 0.02 / x4;              /* L = 0.01 xd^2/x^4 */
                double fel = -0.04 * xd * xd / (x4 * x);
                double h = -0.5 * xd * xd / x4;     /* collision_geometry */
                for (int a = 0; a < DOF; a++) j[a] = k->J[l][0][a] * gh[0] + k->J[l][1][a] * gh[1] + k->J[l][2][a] * gh[2];
                add_leaf(&g, j, Ml, Ml * h, fel, kappa + dot3(gh, cl), dot3(gh, ao + 3 * o), sigma);
            }
            { /* plane leaf: x = (n.p + d)/|n| - r_body */
                double x = dot3(nh, p) + con[3] / nn - rb, xd = dot3(nh, v), j[DOF];
                double s = sw(xd);
                double Ml = 0.2 * s / (x * x);      /* L = 0.1 s xd^2 / x^2 */
                double fel = -0.2 * s * xd * xd / (x * x * x);
                double h = 10.0 * (1.0 / (1.0 + exp(-10.0 * x)) - 1.0) * xd * xd;
                for (int a = 0; a < DOF; a++) j[a] = k->J[l][0][a] * nh[0] + k->J[l][1][a] * nh[1] + k->J[l][2][a] * nh[2];
                add_leaf(&g, j, Ml, Ml * h, fel, dot3(nh, cl), 0.0, sigma);
            }
        }
    }
    for (int i = 0; i < DOF; i++) /* limit leaves: x = q - lo, x = hi - q */
        for (int up = 0; up < 2; up++) {
            double sg = up ? -1.0 : 1.0;
            double x = up ? c->limits[i][1] - q[i] : q[i] - c->limits[i][0], xd = sg * qd[i], j[DOF] = {0};
            double s = sw(xd);
            double Ml = 0.2 * s / x;                /* L = 0.1 s xd^2 / x */
            double fel = -0.1 * s * xd * xd / (x * x);
            double h = -0.1 * xd * xd / x;
            j[i] = sg;
            add_leaf(&g, j, Ml, Ml * h, fel, 0.0, 0.0, sigma);
        }

    /* forced spec = geometry + attractors */
    double Mf[DOF][DOF], ff[DOF];
    memcpy(Mf, g.M, sizeof(Mf));
    memcpy(ff, g.f, sizeof(ff));
    double xpsi_norm;
    { /* sub-goal 0: world -> panda_hand */
        double x[3], w0 = rec[MRFO_W0];
        for (int a = 0; a < 3; a++) x[a] = k->p[7][a] - rec[MRFO_G0 + a];
        double n = sqrt(dot3(x, x));
        xpsi_norm = n;
        double dpsi = 5.0 * w0 * tanh(10.0 * n), m2 = 2.0 * (1.7 * exp(-0.5625 * n * n) + 0.3);
        double t[3];
        for (int a = 0; a < 3; a++) t[a] = m2 * (dpsi * x[a] / n + sigma * k->c[7][a]);
        for (int a = 0; a < DOF; a++) {
            ff[a] += k->J[7][0][a] * t[0] + k->J[7][1][a] * t[1] + k->J[7][2][a] * t[2];
            for (int b = 0; b < DOF; b++)
                Mf[a][b] += m2 * (k->J[7][0][a] * k->J[7][0][b] + k->J[7][1][a] * k->J[7][1][b] + k->J[7][2][a] * k->J[7][2][b]);
        }
    }
    { /* sub-goal 1: R (hand - link7) - x_goal_1 */
        const double* Rg = rec + MRFO_ANG;
        double dlt[3], dc[3], x[3], Jr[3][DOF], cr[3], w1 = rec[MRFO_W1];
        for (int a = 0; a < 3; a++) {
            dlt[a] = k->p[7][a] - k->p[6][a];
            dc[a] = k->c[7][a] - k->c[6][a];
        }
        for (int a = 0; a < 3; a++) {
            x[a] = Rg[3 * a] * dlt[0] + Rg[3 * a + 1] * dlt[1] + Rg[3 * a + 2] * dlt[2] - rec[MRFO_G1 + a];
            cr[a] = Rg[3 * a] * dc[0] + Rg[3 * a + 1] * dc[1] + Rg[3 * a + 2] * dc[2];
            for (int b = 0; b < DOF; b++)
                Jr[a][b] = Rg[3 * a] * (k->J[7][0][b] - k->J[6][0][b]) + Rg[3 * a + 1] * (k->J[7][1][b] - k->J[6][1][b]) +
                           Rg[3 * a + 2] * (k->J[7][2][b] - k->J[6][2][b]);
        }
        double n = sqrt(dot3(x, x));
        double dpsi = 5.0 * w1 * tanh(10.0 * n), m2 = 2.0 * (1.7 * exp(-0.5625 * n * n) + 0.3);
        double t[3];
        for (int a = 0; a < 3; a++) t[a] = m2 * (dpsi * x[a] / n + sigma * cr[a]);
        for (int a = 0; a < DOF; a++) {
            ff[a] += Jr[0][a] * t[0] + Jr[1][a] * t[1] + Jr[2][a] * t[2];
            for (int b = 0; b < DOF; b++) Mf[a][b] += m2 * (Jr[0][a] * Jr[0][b] + Jr[1][a] * Jr[1][b] + Jr[2][a] * Jr[2][b]);
        }
    }
    { /* sub-goal 2: joint 7 -> x_goal_2 (norm_2 of a 1x1 is |x|, gradient sign(x)) */
        double x = q[6] - rec[MRFO_G2], n = fabs(x), w2 = rec[MRFO_W2];
        double sgn = (x > 0) - (x < 0);
        double dpsi = 5.0 * w2 * tanh(10.0 * n), m2 = 2.0 * (1.7 * exp(-0.5625 * n * n) + 0.3);
        ff[6] += m2 * dpsi * sgn;
        Mf[6][6] += m2;
    }

    /* energisation + speed-control damper */
    const double e = c->eps;
    double hg[DOF], hf[DOF];
    if (cholesky_solve(g.M, e, g.f, hg) || cholesky_solve(Mf, e, ff, hf)) {
        /* metric not positive definite (a joint beyond its limit makes the limit-leaf metric 0.2 s/x negative): the
         * reference's symbolic inverse would return garbage here; flag it as NaN so callers can drop the scenario */
        for (int a = 0; a < DOF; a++) action[a] = NAN;
        return 1;
    }
    double num = 0, qMq = 0, qq = 0, qhg = 0, qhf = 0;
    for (int a = 0; a < DOF; a++) {
        num += qd[a] * (g.f[a] - g.fe[a]);
        qq += qd[a] * qd[a];
        qhg += qd[a] * hg[a];
        qhf += qd[a] * hf[a];
        for (int b = 0; b < DOF; b++) qMq += qd[a] * g.M[a][b] * qd[b];
    }
    double a_geom = -num / (e + qMq);
    double s2 = 2.0 * c->exec_scale, den = e + s2 * qq;
    double a_ex0 = -s2 * qhg / den, a_exf = -s2 * qhf / den;
    double eta = 0.5 * (tanh(-0.9 * (1.0 - 1.0 / 2.0) * qq - 0.5) + 1.0);
    double a_ex = eta * a_ex0 + (1.0 - eta) * a_exf;
    double beta = 0.5 * (tanh(-0.5 * (xpsi_norm - 0.02)) + 1.0) * 6.5 + 0.01 + fmax(0.0, a_geom - a_ex);
    double qdd[DOF];
    for (int a = 0; a < DOF; a++) {
        qdd[a] = -hf[a] - (a_ex + beta) * qd[a];
        action[a] = c->mode == 1 ? qd[a] + c->dt * qdd[a] : qdd[a];
    }
    if (diag) {
        double* o = diag;
        memcpy(o, g.M, sizeof(g.M)); o += 49;
        memcpy(o, Mf, sizeof(Mf)); o += 49;
        memcpy(o, g.f, sizeof(g.f)); o += 7;
        memcpy(o, g.fe, sizeof(g.fe)); o += 7;
        memcpy(o, ff, sizeof(ff)); o += 7;
        memcpy(o, qdd, sizeof(qdd));
    }
    return 0;
}

int mrfo_action(const mrfo_config* c, int robot, const double* rec, int S, const double* xo, const double* vo,
                const double* ao, const double* ro, double* action, double* diag) {
    kin_t k;
    kinematics(c->mount[robot], rec + MRFO_Q, rec + MRFO_QD, &k);
    return action_from_kin(c, c->link_mask[robot], &k, rec, S, xo, vo, ao, ro, action, diag);
}

int mrfo_rollout_jointspace(const mrfo_config* c, const double* rec_in, int N, double* qN, double* qdN,
                            double* avg_vel, double* x_ee) {
    return mrfo_rollout_jointspace_static(c, rec_in, N, 0, 0, 0, qN, qdN, avg_vel, x_ee);
}

int mrfo_rollout_jointspace_static(const mrfo_config* c, const double* rec_in, int N, int n_static, const double* xs,
                                   const double* rs, double* qN, double* qdN, double* avg_vel, double* x_ee) {
    const int R = c->n_robots;
    if (n_static < 0 || n_static > MRFO_MAX_STATIC) return 2;
    double rec[MRFO_MAX_ROBOTS][MRFO_ROBOT_IN];
    kin_t kin[MRFO_MAX_ROBOTS];
    double acc[MRFO_MAX_ROBOTS] = {0};
    int rc = 0;
    memcpy(rec, rec_in, sizeof(double) * MRFO_ROBOT_IN * R);
    if (x_ee)
        for (int i = 0; i < R; i++) {
            kinematics(c->mount[i], rec[i] + MRFO_Q, rec[i] + MRFO_QD, &kin[i]);
            memcpy(x_ee + 3 * i, kin[i].p[7], 3 * sizeof(double));
        }
    for (int k = 0; k < N; k++) {
        for (int i = 0; i < R; i++) { /* Phase A: step with the stale velocity, then FK (:191-209) */
            for (int a = 0; a < DOF; a++) rec[i][MRFO_Q + a] += c->dt * rec[i][MRFO_QD + a];
            kinematics(c->mount[i], rec[i] + MRFO_Q, rec[i] + MRFO_QD, &kin[i]);
        }
        double act[MRFO_MAX_ROBOTS][DOF];
        for (int i = 0; i < R; i++) { /* Phase B (:211-249): others ascending j, their collision links */
            enum { SMAX = 8 * (MRFO_MAX_ROBOTS - 1) + MRFO_MAX_STATIC };
            double xo[SMAX][3], vo[SMAX][3], ao[SMAX][3], ro[SMAX];
            int o = 0;
            for (int s = 0; s < n_static; s++, o++) { /* x_obst_s, radius_obst_s of robot i (:319-322): spheres at rest */
                for (int a = 0; a < 3; a++) {
                    xo[o][a] = xs[(i * n_static + s) * 3 + a];
                    vo[o][a] = ao[o][a] = 0.0;
                }
                ro[o] = rs[i * n_static + s];
            }
            for (int j = 0; j < R; j++) {
                if (j == i) continue;
                for (int l = 0; l < 8; l++) {
                    if (!((c->link_mask[j] >> l) & 1)) continue; /* collision_links_nrs[j] */
                    for (int a = 0; a < 3; a++) {
                        xo[o][a] = kin[j].p[l][a];
                        vo[o][a] = c->static_or_dyn ? kin[j].v[l][a] : 0.0;
                        /* a = J qdd + Jdot qdot, qdd = 0, Jdot = jdot_ref_sign * d(J qdot)/dq (utils.py:28,37) */
                        ao[o][a] = c->static_or_dyn ? c->jdot_ref_sign * kin[j].c[l][a] : 0.0;
                    }
                    ro[o] = c->r_robots[j][l];
                    o++;
                }
            }
            rc |= action_from_kin(c, c->link_mask[i], &kin[i], rec[i], o, &xo[0][0], &vo[0][0], &ao[0][0], ro, act[i], 0);
        }
        for (int i = 0; i < R; i++)
            for (int a = 0; a < DOF; a++) {
                rec[i][MRFO_QD + a] = act[i][a];
                acc[i] += act[i][a] * act[i][a];
                if (qN) qN[(i * N + k) * DOF + a] = rec[i][MRFO_Q + a];
                if (qdN) qdN[(i * N + k) * DOF + a] = act[i][a];
            }
    }
    if (avg_vel)
        for (int i = 0; i < R; i++) avg_vel[i] = acc[i] / ((double)N * DOF);
    return rc;
}

int mrfo_rollout_cartesian(const mrfo_config* c, int robot, const double* rec_in, int S, const double* xo_in,
                           const double* vo, const double* ro, int N, double* qN, double* qdN, double* avg_vel) {
    double rec[MRFO_ROBOT_IN], xo[3 * 256], ao[3 * 256], act[DOF], acc = 0;
    int rc = 0;
    if (S > 256) return 2;
    memcpy(rec, rec_in, sizeof(rec));
    memcpy(xo, xo_in, sizeof(double) * 3 * S);
    memset(ao, 0, sizeof(ao));
    for (int k = 0; k < N; k++) {
        rc |= mrfo_action(c, robot, rec, S, xo, vo, ao, ro, act, 0);
        for (int a = 0; a < DOF; a++) {
            if (c->mode == 1) { /* 'vel': the action is the new velocity (forward_planner_Cartesian.py:85-87) */
                rec[MRFO_QD + a] = act[a];
                rec[MRFO_Q + a] += c->dt * act[a];
            } else {            /* 'acc': the action is qdd (:81-84) */
                rec[MRFO_Q + a] += c->dt * rec[MRFO_QD + a] + 0.5 * c->dt * c->dt * act[a];
                rec[MRFO_QD + a] += c->dt * act[a];
            }
            acc += rec[MRFO_QD + a] * rec[MRFO_QD + a];
            if (qN) qN[k * DOF + a] = rec[MRFO_Q + a];
            if (qdN) qdN[k * DOF + a] = rec[MRFO_QD + a];
        }
        for (int o = 0; o < 3 * S; o++) xo[o] += c->dt * vo[o];
    }
    if (avg_vel) *avg_vel = acc / ((double)N * DOF);
    return rc;
}

/* Collision spheres with link-frame offsets (utils.py:87-119 with the transformations of
 * create_simulation_manipulators.py:188-245): x = p_link + R_link t, v_sphere = v_link + omega_link x (R_link t),
 * v_origin = J_link qdot. off [8][n][3]; outputs [8 n][3]. */
void mrfo_spheres(const mrfo_config* c, int robot, const double* q, const double* qd, int n, const double* off,
                  double* x, double* v_origin, double* v_sphere) {
    kin_t k;
    kinematics(c->mount[robot], q, qd, &k);
    for (int l = 0; l < 8; l++)
        for (int s = 0; s < n; s++) {
            const double* t = off + ((size_t)l * n + s) * 3;
            double rt[3], wr[3];
            for (int a = 0; a < 3; a++) rt[a] = k.Rl[l][3 * a] * t[0] + k.Rl[l][3 * a + 1] * t[1] + k.Rl[l][3 * a + 2] * t[2];
            cross(k.wl[l], rt, wr);
            for (int a = 0; a < 3; a++) {
                x[(l * n + s) * 3 + a] = k.p[l][a] + rt[a];
                v_origin[(l * n + s) * 3 + a] = k.v[l][a];
                v_sphere[(l * n + s) * 3 + a] = k.v[l][a] + wr[a];
            }
        }
}

/* Point-mass planner of examples/example_pointmasses_static.py:102-129 / _dynamic.py:102-131: 3 dof (x, y, theta),
 * collision link base_link at fk = (x, y, 0.05) (pointRobot1.urdf:91-113), collision_geometry "-2/x xdot^2",
 * collision_finsler "1/x^2 (1 - heaviside(xdot)) xdot^2", one 2-D attractor on (x, y), no limits, mode 'acc'.
 * static spheres xs [Ss][3], rs [Ss]; dynamic spheres (2-D) xd, vd, ad [Sd][2], rd [Sd]. */
int mrfo_point_action(const mrfo_config* c, const double* q, const double* qd, const double* goal, double w_goal,
                      double r_body, int Ss, const double* xs, const double* rs, int Sd, const double* xd,
                      const double* vd, const double* ad, const double* rd, double* action) {
    const double sigma = c->jdot_sign, e = c->eps;
    double M[2][2] = {{0.2, 0.0}, {0.0, 0.2}}, f[2] = {0, 0}, num = 0; /* theta decouples: M33 = 0.2, f3 = 0 */
    for (int o = 0; o < Ss + Sd; o++) {
        const int dyn = o >= Ss;
        const int k = dyn ? o - Ss : o;
        double d[3], w[3], ao[3] = {0, 0, 0}, rho;
        if (!dyn) {
            d[0] = q[0] - xs[3 * k]; d[1] = q[1] - xs[3 * k + 1]; d[2] = 0.05 - xs[3 * k + 2];
            w[0] = qd[0]; w[1] = qd[1]; w[2] = 0.0;
            rho = rs[k] + r_body;
        } else {
            d[0] = q[0] - xd[2 * k]; d[1] = q[1] - xd[2 * k + 1]; d[2] = 0.0;
            w[0] = qd[0] - vd[2 * k]; w[1] = qd[1] - vd[2 * k + 1]; w[2] = 0.0;
            ao[0] = ad[2 * k]; ao[1] = ad[2 * k + 1];
            rho = rd[k] + r_body;
        }
        double n = sqrt(dot3(d, d)), x = n / rho - 1.0, g[3] = {d[0] / (n * rho), d[1] / (n * rho), d[2] / (n * rho)};
        double xdot = dot3(g, w), dw = dot3(d, w);
        double kappa = (dot3(w, w) - dw * dw / (n * n)) / (n * rho);
        double s = xdot < 0 ? 1.0 : (xdot > 0 ? 0.0 : 0.5);     /* 1 - heaviside(xdot), heaviside(0) = 0.5 */
        double Ml = 2.0 * s / (x * x), fel = -2.0 * s * xdot * xdot / (x * x * x), h = -2.0 / x * xdot * xdot;
        double fl = Ml * h, acc_o = dot3(g, ao);
        double fq = fl + Ml * (sigma * kappa - acc_o), feq = fel + Ml * (kappa - acc_o);
        double gv = g[0] * qd[0] + g[1] * qd[1]; /* J^T g . qdot */
        for (int a = 0; a < 2; a++) {
            f[a] += g[a] * fq;
            for (int b = 0; b < 2; b++) M[a][b] += Ml * g[a] * g[b];
        }
        num += gv * (fq - feq);
    }
    double Mf[2][2] = {{M[0][0], M[0][1]}, {M[1][0], M[1][1]}}, ff[2] = {f[0], f[1]};
    double xg[2] = {q[0] - goal[0], q[1] - goal[1]}, n = sqrt(xg[0] * xg[0] + xg[1] * xg[1]);
    double dpsi = 5.0 * w_goal * tanh(10.0 * n), m2 = 2.0 * (1.7 * exp(-0.5625 * n * n) + 0.3);
    for (int a = 0; a < 2; a++) {
        ff[a] += m2 * dpsi * xg[a] / n;
        Mf[a][a] += m2;
    }
    double hg[3], hf[3];
    for (int sys = 0; sys < 2; sys++) {
        double (*A)[2] = sys ? Mf : M;
        double* b = sys ? ff : f;
        double* hh = sys ? hf : hg;
        double a00 = A[0][0] + e, a11 = A[1][1] + e, a01 = A[0][1], det = a00 * a11 - a01 * a01;
        hh[0] = (a11 * b[0] - a01 * b[1]) / det;
        hh[1] = (a00 * b[1] - a01 * b[0]) / det;
        hh[2] = 0.0;
    }
    double qMq = 0.2 * qd[2] * qd[2], qq = 0, qhg = 0, qhf = 0;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) qMq += qd[a] * M[a][b] * qd[b];
    for (int a = 0; a < 3; a++) {
        qq += qd[a] * qd[a];
        qhg += qd[a] * hg[a];
        qhf += qd[a] * hf[a];
    }
    double a_geom = -num / (e + qMq), s2 = 2.0 * c->exec_scale, den = e + s2 * qq;
    double a_ex0 = -s2 * qhg / den, a_exf = -s2 * qhf / den;
    double eta = 0.5 * (tanh(-0.9 * (1.0 - 1.0 / 2.0) * qq - 0.5) + 1.0);
    double a_ex = eta * a_ex0 + (1.0 - eta) * a_exf;
    double beta = 0.5 * (tanh(-0.5 * (n - 0.02)) + 1.0) * 6.5 + 0.01 + fmax(0.0, a_geom - a_ex);
    for (int a = 0; a < 3; a++) action[a] = -hf[a] - (a_ex + beta) * qd[a]; /* mode 'acc' */
    return 0;
}

int mrfo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Host cores this process may run on, whatever OMP_NUM_THREADS says (a launcher such as torchrun exports
 * OMP_NUM_THREADS=1 to every rank; the CPU arm of bench.py sets its thread count explicitly from this). */
int mrfo_hw_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int mrfo_rollout_jointspace_batch(const mrfo_config* c, const double* rec, long batch, int N, double* qN,
                                  double* qdN, double* avg_vel, double* x_ee, int n_threads) {
    const int R = c->n_robots;
    int rc = 0;
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : mrfo_max_threads()) reduction(| : rc)
    for (long b = 0; b < batch; b++)
        rc |= mrfo_rollout_jointspace(c, rec + b * R * MRFO_ROBOT_IN, N, qN ? qN + b * R * N * DOF : 0,
                                      qdN ? qdN + b * R * N * DOF : 0, avg_vel ? avg_vel + b * R : 0,
                                      x_ee ? x_ee + b * R * 3 : 0);
    return rc;
}

int mrfo_action_batch(const mrfo_config* c, int robot, const double* rec, long batch, int S, const double* xo,
                      const double* vo, const double* ao, const double* ro, double* action, int n_threads) {
    int rc = 0;
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads > 0 ? n_threads : mrfo_max_threads()) reduction(| : rc)
    for (long b = 0; b < batch; b++)
        rc |= mrfo_action(c, robot, rec + b * MRFO_ROBOT_IN, S, xo + b * 3 * S, vo + b * 3 * S, ao + b * 3 * S,
                          ro + b * S, action + b * DOF, 0);
    return rc;
}

/* RF-CV control-step rollout, batched: per scenario the constant-velocity goal estimate of robot `est_robot`
 * (example_pandas_Jointspace.py:346-348: goal_1 = x_ee + est_h * v_ee, with v_ee the first Jacobian column unless
 * use_jqd) followed by the coupled rollout -- the whole per-scenario path inside one OpenMP loop, nothing left to a
 * Python loop.  rec is not modified; goal_est (nullable) [batch][3] receives the estimate. */
int mrfo_rollout_rfcv_batch(const mrfo_config* c, const double* rec, long batch, int N, int est_robot, double est_h,
                            int use_jqd, double* qN, double* qdN, double* avg_vel, double* x_ee, double* goal_est,
                            int n_threads) {
    const int R = c->n_robots;
    int rc = 0;
    if (est_robot >= R) est_robot = -1;
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : mrfo_max_threads()) reduction(| : rc)
    for (long b = 0; b < batch; b++) {
        double loc[MRFO_MAX_ROBOTS * MRFO_ROBOT_IN];
        memcpy(loc, rec + b * R * MRFO_ROBOT_IN, sizeof(double) * (size_t)R * MRFO_ROBOT_IN);
        if (est_robot >= 0) {
            double* rr = loc + est_robot * MRFO_ROBOT_IN;
            double x[3], v[3];
            mrfo_endeffector(c, est_robot, rr + MRFO_Q, rr + MRFO_QD, use_jqd, x, v);
            for (int k = 0; k < 3; k++) rr[MRFO_G0 + k] = x[k] + est_h * v[k];
            if (goal_est)
                for (int k = 0; k < 3; k++) goal_est[b * 3 + k] = rr[MRFO_G0 + k];
        }
        rc |= mrfo_rollout_jointspace(c, loc, N, qN ? qN + b * R * N * DOF : 0, qdN ? qdN + b * R * N * DOF : 0,
                                      avg_vel ? avg_vel + b * R : 0, x_ee ? x_ee + b * R * 3 : 0);
    }
    return rc;
}
