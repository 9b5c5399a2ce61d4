/* oracle/oracle_capi.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * C surface of the CPU oracle (fp64 restatement of compute_grf -> ConvexMpc ->
 * OSQP).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  PARITY UNPINNED: see mpc_oracle.hpp. */
#ifndef ORACLE_CAPI_H
#define ORACLE_CAPI_H

#include "../include/mpc_b200.h" /* record and config types only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleResult {
  double grf[12]; /* body-frame first-step GRF (or x[0..12) for oracle_osqp_solve_mpc) */
  int32_t status;
  int32_t iters;
  int32_t rho_updates;
  int32_t pad;
  double pri_res;
  double dua_res;
} OracleResult;

/* ConvexMpc.cpp:110-245 + A1RobotControl.cpp:452-518 for one state. */
int oracle_mpc_build_qp(const MpcConfig *cfg, const MpcStateIn *state, double *P, double *q,
                        double *l, double *u);
int oracle_mpc_build_intermediates(const MpcConfig *cfg, const MpcStateIn *state, double *A_d,
                                   double *B_d, double *A_qp, double *B_qp, double *x0,
                                   double *x_ref);
/* A1RobotControl.cpp:446-561 for n states, OpenMP one problem per thread.
 * solutions (optional): n x 12H full primal solutions, world frame. */
int oracle_mpc_compute_grf(const MpcConfig *cfg, const MpcStateIn *states, int32_t n,
                           OracleResult *out, double *solutions, int32_t threads);
/* Same code instantiated in float: an arithmetic model of an fp32 device path. */
int oracle_mpc_compute_grf_f32(const MpcConfig *cfg, const MpcStateIn *states, int32_t n,
                               OracleResult *out, double *solutions, int32_t threads);
/* ConvexMpc::calculate_qp_mats from caller-written A_mat_d / B_mat_d_list. */
int oracle_qp_mats_from_model(const MpcConfig *cfg, const double *A_mat_d,
                              const double *B_mat_d_list, const double *mpc_states,
                              const double *mpc_states_d, const int32_t *contacts,
                              double *hessian, double *gradient, double *lb, double *ub);
/* OSQP restatement on a caller-supplied dense MPC QP (friction-pyramid A). */
int oracle_osqp_solve_mpc(const MpcConfig *cfg, const double *P, const double *q, const double *l,
                          const double *u, double *x, double *y, OracleResult *info);
/* A1RobotControl.cpp:377-444. */
int oracle_balance_build_qp(const BalanceConfig *cfg, const BalanceStateIn *state, double *P,
                            double *q, double *l, double *u);
int oracle_balance_compute_grf(const BalanceConfig *cfg, const BalanceStateIn *states, int32_t n,
                               OracleResult *out, int32_t threads);
/* A1RobotControl.cpp:522-540 kept alive over `ticks` control ticks (states tick-major). */
int oracle_mpc_stream(const MpcConfig *cfg, const MpcStateIn *states, int32_t n, int32_t ticks,
                      OracleResult *out, int32_t threads);
/* A1RobotControl.cpp:289-319 from given body-frame GRFs. */
int oracle_torque_map(const float *state_words, int32_t state_stride, int32_t contact_offset,
                      const MpcTorqueIn *tin, const double *grf, int32_t n, double *joint_torques,
                      int32_t *nan_mask);
/* GazeboA1ROS.cpp:262-288,306 + A1BasicEKF.cpp:54-164 + A1RobotControl.cpp:335-376,566-582 over
 * `ticks` consecutive sensor batches (tick-major), persistent estimator per robot. */
int oracle_prep_stream(const PrepConfig *cfg, const RobotSensorIn *sensors, int32_t n, int32_t ticks,
                       MpcStateIn *states, MpcTorqueIn *tin, RobotPrepOut *extras);
int oracle_leg_fk_jac(const double *rho_fix, const double *q, double *p, double *J);
/* SURVEY.md 8f row 4 (flags in cfg; gait may be NULL when gait_aware = 0). */
int oracle_mpc_build_qp_ext(const MpcConfig *cfg, const MpcStateIn *state, const MpcGaitIn *gait, double *P,
                            double *q, double *l, double *u);
int oracle_mpc_compute_grf_ext(const MpcConfig *cfg, const MpcStateIn *states, const MpcGaitIn *gait, int32_t n,
                               OracleResult *out, int32_t threads);
int oracle_discretize_exact(const MpcConfig *cfg, const MpcStateIn *state, double *A_d, double *B_d);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
