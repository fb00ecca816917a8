/* Oracle O2 -- closed-form float64 CPU restatement of the multi-robot fabric hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker / CPU baseline.  The product (libmrf_b200.so) never links it.
 *
 * PARITY UNPINNED: fabrics==0.9.5 / forwardkinematics==1.2.3 / casadi==3.5.5 are not vendored in
 * /root/reference and cannot be installed offline; the reference's tests hold no golden vectors.
 * This restates the published fabrics algorithm in closed form and is pinned against oracle O1
 * (oracle/o1_fabrics.py, autodiff, fabrics-structured) via tests/golden/.
 */
#ifndef MRF_ORACLE_H
#define MRF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define MRFO_MAX_ROBOTS 8
#define MRFO_DOF 7
#define MRFO_NLINKS 8          /* panda_link1..8 (parameters_manipulators.py:25-26) */
#define MRFO_ROBOT_IN 44       /* packed per-robot rollout input record, see below */
#define MRFO_MAX_STATIC 16     /* static spheres per robot in a coupled rollout */

/* Per-robot record (doubles), the arguments of get_velocity_rollouts
 * (forward_planner_Jointspace.py:303-329) in a fixed order:
 *  [0..6] q  [7..13] qdot  [14..16] x_goal_0  [17] weight_goal_0  [18..20] x_goal_1  [21] weight_goal_1
 *  [22] x_goal_2  [23] weight_goal_2  [24..32] angle_goal_1 (row-major 3x3)  [33..36] constraint_0
 *  [37..42] radius_body_panda_link3..8  [43] pad */
enum { MRFO_Q = 0, MRFO_QD = 7, MRFO_G0 = 14, MRFO_W0 = 17, MRFO_G1 = 18, MRFO_W1 = 21, MRFO_G2 = 22,
       MRFO_W2 = 23, MRFO_ANG = 24, MRFO_CON = 33, MRFO_RB = 37 };

typedef struct {
    int n_robots;
    int mode;                 /* 0 = 'acc', 1 = 'vel' (concretize(mode, time_step)) */
    int static_or_dyn;        /* STATIC_OR_DYN_FABRICS: 0 zeroes other robots' v and a in rollouts */
    int has_collision_links;  /* 0 = grasp planner (collision_links_nr=[]): no sphere/plane leaves */
    double dt;                /* planner time_step and rollout dt (parameters_manipulators.py:8) */
    double eps;               /* fabrics eps = 1e-6 */
    double jdot_sign;         /* fabrics DifferentialMap Jdot_sign (-1) */
    double jdot_ref_sign;     /* reference utils.py:28 Jdot_sign (-1) for published accelerations */
    double exec_scale;        /* ExecutionLagrangian = exec_scale * qdot.qdot */
    double mount[MRFO_MAX_ROBOTS][16];     /* row-major 4x4 (example_pandas_Jointspace.py:108-118) */
    double limits[MRFO_DOF][2];            /* example_pandas_Jointspace.py:97-105 */
    double r_robots[MRFO_MAX_ROBOTS][MRFO_NLINKS]; /* other-robot sphere radii, compile-time in the
                                                      reference graph (forward_planner_Jointspace.py:221) */
    int link_mask[MRFO_MAX_ROBOTS];               /* bit l-1: panda_link l in collision_links_nr of the robot
                                                      (example_pandas_Jointspace.py:64,91-96); default 0xFF */
} mrfo_config;

void mrfo_config_default(mrfo_config* c, int n_robots);

/* Kinematics of robot `robot`: link origins x[8][3], v = J qdot, c = (+) d(J qdot)/dq qdot, Jacobians. */
void mrfo_kinematics(const mrfo_config* c, int robot, const double* q, const double* qd,
                     double x[8][3], double v[8][3], double cdd[8][3], double J[8][3][7]);

/* One fabric action (planner._funs._function, no small-action clamp).  rec = per-robot record;
 * obstacles: S spheres, xo/vo/ao [S][3], ro [S].  diag (nullable, 7*7*2+7*3+7 doubles):
 * M_g, M_f, f_g, fe_g, f_f, qdd. Returns 0, or 1 if the Cholesky factorisation failed. */
int mrfo_action(const mrfo_config* c, int robot, const double* rec, int S, const double* xo,
                const double* vo, const double* ao, const double* ro, double* action, double* diag);

/* ForwardFabricsPlanner coupled rollout (forward_planner_Jointspace.py:190-249), one scenario.
 * rec [R][44]; outputs (nullable): qN, qdN [R][N][7]; avg_vel [R]; x_ee [R][3] = hand at the input q. */
int mrfo_rollout_jointspace(const mrfo_config* c, const double* rec, int N, double* qN, double* qdN,
                            double* avg_vel, double* x_ee);

/* The same with n_static static spheres per robot (x_obst_s, radius_obst_s of the rollout planners,
 * forward_planner_Jointspace.py:319-322): xs [R][n_static][3], rs [R][n_static]. */
int mrfo_rollout_jointspace_static(const mrfo_config* c, const double* rec, int N, int n_static, const double* xs,
                                   const double* rs, double* qN, double* qdN, double* avg_vel, double* x_ee);

/* FabricsRollouts decoupled rollout (forward_planner_Cartesian.py:421-458), one robot. */
int mrfo_rollout_cartesian(const mrfo_config* c, int robot, const double* rec, int S, const double* xo,
                           const double* vo, const double* ro, int N, double* qN, double* qdN, double* avg_vel);

/* Batched (OpenMP over scenarios) versions; arrays are [batch] x the single-scenario layout. */
int mrfo_rollout_jointspace_batch(const mrfo_config* c, const double* rec, long batch, int N, double* qN,
                                  double* qdN, double* avg_vel, double* x_ee, int n_threads);
int mrfo_action_batch(const mrfo_config* c, int robot, const double* rec, long batch, int S, const double* xo,
                      const double* vo, const double* ao, const double* ro, double* action, int n_threads);

/* RF-CV goal estimate (example_pandas_Jointspace.py:236-238,328-329,346-348): hand position and the
 * FIRST COLUMN of its Jacobian (the reference's "v_ee", quirk Q2), or J qdot if use_jqd != 0
 * (example_pandas_cartesian.py / utils.py:131). */
void mrfo_endeffector(const mrfo_config* c, int robot, const double* q, const double* qd, int use_jqd,
                      double x_ee[3], double v_ee[3]);

/* Collision spheres with link-frame offsets off[8][n][3] (utils.py:87-119, create_simulation_manipulators.py:188-245):
 * x, v_origin (= J_link qdot), v_sphere (= J_sphere qdot), each [8 n][3]. */
void mrfo_spheres(const mrfo_config* c, int robot, const double* q, const double* qd, int n, const double* off,
                  double* x, double* v_origin, double* v_sphere);

/* Point-mass planner (examples/example_pointmasses_static.py:102-129, _dynamic.py:102-131), mode 'acc'. */
int mrfo_point_action(const mrfo_config* c, const double* q, const double* qd, const double* goal, double w_goal,
                      double r_body, int Ss, const double* xs, const double* rs, int Sd, const double* xd,
                      const double* vd, const double* ad, const double* rd, double* action);

/* RF-CV control-step rollout: goal estimate of robot est_robot (-1: none; example_pandas_Jointspace.py:346-348) +
 * coupled rollout, per scenario inside one OpenMP loop.  goal_est (nullable) [batch][3]. */
int mrfo_rollout_rfcv_batch(const mrfo_config* c, const double* rec, long batch, int N, int est_robot, double est_h,
                            int use_jqd, double* qN, double* qdN, double* avg_vel, double* x_ee, double* goal_est,
                            int n_threads);

int mrfo_max_threads(void);
int mrfo_hw_threads(void); /* cores available to the process, ignoring OMP_NUM_THREADS */

#ifdef __cplusplus
}
#endif
#endif
