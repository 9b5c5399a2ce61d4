/*
 * mpc_b200.h -- C ABI of the B200-native batched convex-MPC ground-reaction-force
 * (GRF) engine.
 *
 * This header is the drop-in boundary for ONE hot path of
 * zerenluo123/Go1-QP-MPC-Controller:
 *
 *   A1RobotControl::compute_grf (MPC branch and stance-QP branch)
 *       -> ConvexMpc::{reset, calculate_A_mat_c, calculate_B_mat_c,
 *                      state_space_discretization, calculate_qp_mats}
 *       -> OsqpEigen::Solver::{initSolver, solve, getSolution}
 *
 * The reference has no FFI: the path is ordinary C++ member calls
 * (src/a1_cpp/src/ConvexMpc.h:22-35, src/a1_cpp/src/A1RobotControl.h:44).  The
 * entry points below are what a binding for that path would bind; every one
 * cites the reference code it replaces.  All functions are extern "C", take
 * plain pointers and sizes, return int (0 = ok, <0 = MpcError) and never throw.
 * There is NO CPU fallback behind this ABI: without a CUDA device every compute
 * entry point fails with MPC_ERR_NO_DEVICE.
 *
 * Conventions
 *   - leg order FL, FR, RL, RR everywhere (A1CtrlStates.h:44-47, :400).
 *   - 3x3 matrices are row-major float[9]; foot positions are leg-major xyz.
 *   - caller owns all host buffers; the engine owns its device memory.
 *   - one engine per host thread; one CUDA stream per engine.
 *   - functions are synchronous unless suffixed _async.  Exception, stated here once: the loaders
 *     (mpc_load_states, mpc_set_torque_inputs, mpc_set_gait_inputs, mpc_prepare_states,
 *     balance_load_states) ENQUEUE their host-to-device copy on the engine stream and return; with a
 *     pinned host buffer that copy is truly asynchronous, so the caller must not rewrite the buffer
 *     before the next synchronising call on the engine (mpc_get_results, mpc_get_torques,
 *     mpc_synchronize, any non-_async solve).  The one-call paths (mpc_compute_grf_batch,
 *     mpc_stream_step, balance_qp_solve) end in such a call.
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compile-time shapes of the reference (A1Params.h:26-34). */
#define MPC_STATE_DIM 13
#define MPC_NUM_DOF 12
#define MPC_CONSTRAINT_DIM 20
#define MPC_NUM_LEG 4
#define MPC_HORIZON_DEFAULT 10 /* PLAN_HORIZON, A1Params.h:26 */
#define MPC_HORIZON_MAX 30
#define MPC_INFTY 1e30 /* OsqpEigen::INFTY == OSQP_INFTY (ConvexMpc.cpp:229-237) */

typedef enum MpcError {
  MPC_OK = 0,
  MPC_ERR_INVALID = -1,   /* bad argument / config */
  MPC_ERR_NO_DEVICE = -2, /* no CUDA device: there is no CPU fallback */
  MPC_ERR_CUDA = -3,      /* CUDA runtime error; text in mpc_last_error */
  MPC_ERR_STATE = -4,     /* call order violated (e.g. solve before build) */
  MPC_ERR_UNSUPPORTED = -5
} MpcError;

/* OSQP status codes, kept so callers of solver.solve()/getStatus see the same
 * values (osqp 0.6.x constants.h; the reference drops the status,
 * A1RobotControl.cpp:431,540). */
#define MPC_STATUS_SOLVED 1
#define MPC_STATUS_MAX_ITER_REACHED (-2)
/* -3 / -4 are OSQP's certificates of infeasibility.  The DEVICE solvers never produce them: every QP of
 * this path is feasible (f = 0 satisfies all rows) and strictly convex (P >= 2 r I), so the certificates
 * cannot fire (the oracle evaluates them and the tests assert that they never do); eps_prim_inf /
 * eps_dual_inf are therefore accepted and unused on the device.  mpc_solve_qp, the one entry that takes
 * caller bounds, refuses lb > ub with MPC_ERR_INVALID like osqp_setup does; a caller QP that is infeasible
 * in a subtler way ends as MPC_STATUS_MAX_ITER_REACHED where OSQP would report -3 / -4. */
#define MPC_STATUS_PRIMAL_INFEASIBLE (-3)
#define MPC_STATUS_DUAL_INFEASIBLE (-4)
#define MPC_STATUS_UNSOLVED (-10)
#define MPC_STATUS_INTERNAL_ERROR (-100) /* device-side synchronisation timed out: a bug, never expected */

/* One robot state for the MPC branch: the A1CtrlStates subset read by
 * compute_grf (A1RobotControl.cpp:452-488, :498-514; fields A1CtrlStates.h:
 * 348-352, 373-379, 400, 412).  48 x fp32 = 192 B, 16 B aligned so a warp
 * reads records with 128-bit loads. */
typedef struct MpcStateIn {
  float euler[3];        /* root_euler (roll, pitch, yaw), ZYX (Utils.cpp:7-33) */
  float pos[3];          /* root_pos */
  float ang_vel[3];      /* root_ang_vel (world) */
  float lin_vel[3];      /* root_lin_vel (world) */
  float euler_d[3];      /* root_euler_d; [2] is never read by the reference */
  float pos_d_z;         /* root_pos_d[2] */
  float lin_vel_d[3];    /* root_lin_vel_d, BODY frame (rotated at :470) */
  float ang_vel_d[3];    /* root_ang_vel_d */
  float rot_mat[9];      /* root_rot_mat, row-major */
  float foot_pos_abs[12]; /* foot_pos_abs, leg-major xyz */
  float contacts[4];     /* contacts[i] as 0.0f / 1.0f */
  float pad;
} MpcStateIn;

/* One robot state for the stance-balance QP branch (A1RobotControl.cpp:377-444).
 * 64 x fp32 = 256 B. */
typedef struct BalanceStateIn {
  float euler[3];
  float pos[3];
  float ang_vel[3];
  float lin_vel[3];
  float euler_d[3];
  float pos_d[3];
  float lin_vel_d[3]; /* body frame */
  float ang_vel_d[3];
  float rot_mat[9];   /* root_rot_mat */
  float rot_mat_z[9]; /* root_rot_mat_z (yaw only) */
  float foot_pos_abs[12];
  float contacts[4];
  float pad[6];
} BalanceStateIn;

/* One result: first-step GRF in the BODY frame (Rᵀ·f, A1RobotControl.cpp:
 * 439-444, :558-561), 3 floats per leg.  On a NaN solution the reference leaves
 * the force untouched (:559); the engine writes zeros and flags status.
 * 64 B. */
typedef struct MpcResult {
  float grf[12];
  int32_t status;      /* MPC_STATUS_* */
  int32_t iters;       /* ADMM iterations run */
  int32_t rho_updates; /* refactorisations after the first */
  float pri_res;       /* unscaled primal residual at exit */
} MpcResult;

/* Inputs of the torque map (compute_joint_torques, A1RobotControl.cpp:289-319)
 * that compute_grf does not already read; 256 B. */
typedef struct MpcTorqueIn {
  float j_foot[36];          /* the four 3x3 diagonal blocks of state.j_foot, leg-major, each ROW-major */
  float foot_forces_kin[12]; /* swing-leg PD force, robot frame, 3 per leg (A1RobotControl.cpp:254-255) */
  float km_foot[3];          /* A1CtrlStates.h:122 */
  float torques_gravity[12]; /* A1CtrlStates.h:129 */
  float pad;
} MpcTorqueIn;

/* joint_torques of one robot; bit i of nan_mask set = component i was NaN and the
 * caller keeps its previous value (A1RobotControl.cpp:314-317); 64 B. */
typedef struct MpcTorqueOut {
  float joint_torques[12];
  int32_t nan_mask;
  int32_t pad[3];
} MpcTorqueOut;

/* What the sensor callbacks and the joystick/gait code leave in A1CtrlStates before a control
 * tick (GazeboA1ROS.cpp:216-308, HardwareA1ROS.cpp:270-350): the input of the on-device state
 * preparation.  384 B. */
typedef struct RobotSensorIn {
  float joint_pos[12];              /* FL, FR, RL, RR x (hip, thigh, calf) */
  float joint_vel[12];
  float root_quat[4];               /* w, x, y, z */
  float imu_acc[3];
  float imu_ang_vel[3];             /* body frame */
  float foot_force[4];              /* contact sensors, N */
  float root_pos[3];                /* odometry; used when the estimator is off or on its first tick */
  float root_lin_vel[3];
  float root_euler_d[3];            /* commands */
  float root_pos_d_z;
  float root_lin_vel_d[3];
  float root_ang_vel_d[3];
  float contacts[4];                /* gait scheduler / early-contact logic, 0 or 1 */
  float foot_pos_recent_contact[12];/* 3 per leg, filtered by the swing-leg code (A1RobotControl.cpp:271-279) */
  float foot_forces_kin[12];        /* swing PD force, 3 per leg */
  float movement_mode;              /* 0 stand, 1 walk (A1BasicEKF.cpp:79-86) */
  float dt;                         /* estimator step */
  float pad[12];
} RobotSensorIn;

/* Derived quantities of the state preparation that do not travel in MpcStateIn; 512 B. */
typedef struct RobotPrepOut {
  float root_euler[3];
  float root_rot_mat[9];
  float root_ang_vel[3];
  float foot_pos_rel[12];           /* 3 per leg */
  float foot_vel_rel[12];
  float foot_pos_abs[12];
  float foot_vel_abs[12];
  float foot_pos_world[12];
  float foot_vel_world[12];
  float estimated_root_pos[3];
  float estimated_root_vel[3];
  float estimated_contacts[4];
  float terrain_pitch_angle;
  float root_euler_d_pitch;         /* root_euler_d[1] after terrain adaptation */
  float pad[29];
} RobotPrepOut;

/* Engine-wide constants of the state preparation. */
typedef struct PrepConfig {
  double rho_fix[20];        /* per leg: leg_offset_x, leg_offset_y, motor_offset, upper, lower (GazeboA1ROS.cpp:76-93) */
  double km_foot[3];         /* A1CtrlStates.h:122 */
  double torques_gravity[12];/* A1CtrlStates.h:129 */
  int32_t use_estimator;     /* 1: A1BasicEKF writes root_pos / root_lin_vel (A1BasicEKF.cpp:160-163) */
  int32_t assume_flat_ground;/* A1BasicEKF.cpp:43-53 */
  int32_t use_terrain_adapt; /* A1RobotControl.cpp:358-364 */
  int32_t pad;
} PrepConfig;

/* OSQP settings (osqp 0.6.x names).  The reference leaves all but verbose and
 * warm_start at library defaults (A1RobotControl.cpp:523-524). */
typedef struct MpcSolverSettings {
  double rho;
  double sigma;
  double alpha;
  double eps_abs;
  double eps_rel;
  double eps_prim_inf;
  double eps_dual_inf;
  int32_t max_iter;
  int32_t check_termination;
  int32_t scaling; /* Ruiz iterations */
  int32_t adaptive_rho;
  int32_t adaptive_rho_interval; /* pinned: the library default is wall-clock dependent */
  double adaptive_rho_tolerance;
} MpcSolverSettings;

/* Engine-wide constants: what ConvexMpc's constructor and A1CtrlStates hold
 * once per robot model (ConvexMpc.cpp:7-68; A1CtrlStates.h:359-367). */
typedef struct MpcConfig {
  int32_t horizon;      /* PLAN_HORIZON; 10 or 30 */
  int32_t reserved0;
  double dt;            /* mpc_dt, A1RobotControl.cpp:462 */
  double mu;            /* ConvexMpc.cpp:8 */
  double fz_min;        /* ConvexMpc.cpp:223 */
  double fz_max;        /* ConvexMpc.cpp:224 */
  double mass;          /* robot_mass */
  double inertia[9];    /* a1_trunk_inertia, row-major */
  double q_weights[13]; /* state.q_weights */
  double r_weights[12]; /* state.r_weights */
  MpcSolverSettings osqp;
  /* behaviours the reference's authors sketched but left switched off.  Each
   * CHANGES RESULTS, defaults 0 = the reference's behaviour; H = 10 only. */
  int32_t exact_discretization; /* A_d, B_d from the matrix exponential the reference commented out
                                   (ConvexMpc.cpp:149); A_c is nilpotent, so the series is exact:
                                   A_d = I + A dt + A^2 dt^2/2, B_d = (dt I + A dt^2/2) B_c */
  int32_t foot_drift;           /* per-step B_d: lever arms r_i = r_0 - i dt v_d (world), the update
                                   commented out at A1RobotControl.cpp:504-507 */
  int32_t gait_aware;           /* per-step contacts from the gait counters (mpc_set_gait_inputs)
                                   instead of today's contacts replicated (ConvexMpc.cpp:242-245) */
  int32_t structured_solver;    /* how K x = r is solved inside the ADMM (same iterates up to rounding):
                                   0 automatic: the fused wrench-space kernels, build and solve in one launch,
                                     no Hessian in memory -- H = 10: 60 x 60 Woodbury core in registers
                                     (wrench_tile_kernel.cuh, 4 x 8 tiles; wrench_kernel.cuh, half rows, with
                                     MPC_WRENCH_TILE=0 in the environment); H = 30: the core solved as a 12-state / 6-input
                                     Riccati recursion (wrench_riccati_kernel.cuh); H = 30 with
                                     exact_discretization: as 1;
                                   1 Riccati recursion on the dense build (riccati_kernel.cuh; at H = 10 cold
                                     solves only, warm-started solves stay dense);
                                   2 dense: qp_build_kernel + admm_solve_kernel, K^-1 (120 x 120) in registers
                                     (for H = 30: K^-1 in a per-CTA L2 workspace, 7x slower);
                                   3 wrench-space, explicitly (MPC_ERR_INVALID for H = 30 with
                                     exact_discretization) */
} MpcConfig;

/* Gait scheduler state of one robot (A1CtrlStates.h:24-28,103; A1RobotControl.cpp:156-164),
 * for gait_aware = 1.  Step 0 of the horizon keeps MpcStateIn.contacts; step i >= 1 is in
 * contact iff fmod(gait_counter + i ticks_per_step gait_counter_speed, counter_per_gait)
 * <= counter_per_swing.  48 B. */
typedef struct MpcGaitIn {
  float gait_counter[4];
  float gait_counter_speed[4];
  float counter_per_gait;
  float counter_per_swing;
  float ticks_per_step; /* control ticks per MPC step (mpc_dt / control dt) */
  float pad;
} MpcGaitIn;

/* Constants of the stance-balance QP (A1RobotControl.cpp:11-15) plus the PD
 * gains it reads from A1CtrlStates (A1CtrlStates.h:429-432). */
typedef struct BalanceConfig {
  double Q[6];  /* 1,1,1,400,400,100 */
  double R;     /* 1e-3 */
  double mu;    /* 0.7 */
  double F_min; /* 0 */
  double F_max; /* 180 */
  double mass;
  double kp_linear[3];
  double kd_linear[3];
  double kp_angular[3];
  double kd_angular[3];
  MpcSolverSettings osqp;
} BalanceConfig;

typedef struct MpcEngine MpcEngine;

/* ---- configuration ------------------------------------------------------ */

/* OSQP 0.6.x library defaults, i.e. what the reference runs with. */
int mpc_settings_osqp_default(MpcSolverSettings *s);
/* BASELINE.json benchmark settings: eps_abs = eps_rel = 1e-5, interval 50. */
int mpc_settings_benchmark(MpcSolverSettings *s);
/* config/gazebo_a1_mpc.yaml weights and mass, H=10, dt 0.0025, mu 0.3,
 * fz in [0,180], benchmark solver settings. */
int mpc_config_default(MpcConfig *cfg);
/* config/hardware_a1_mpc.yaml weights and mass (the well-conditioned set). */
int mpc_config_hardware(MpcConfig *cfg);
/* A1RobotControl.cpp:11-15 constants + config/gazebo_a1_qp.yaml gains. */
int balance_config_default(BalanceConfig *cfg);

/* ---- synthetic inputs (SURVEY.md 8d generator; host only, no device) ----- */

/* Fills out[0..n) with the states first_index .. first_index+n-1 of the
 * counter-based stream `seed`.  Deterministic and order independent. */
int mpc_generate_states(uint64_t seed, uint64_t first_index, int32_t n, MpcStateIn *out);
int balance_generate_states(uint64_t seed, uint64_t first_index, int32_t n, BalanceStateIn *out);
/* The same robots `tick` control periods (2.5 ms) later: velocities relax to the commanded
 * ones, the pose integrates them, trot pairs swap every 48 ticks.  tick 0 = mpc_generate_states. */
int mpc_generate_stream_states(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick,
                               MpcStateIn *out);

/* ---- engine life cycle --------------------------------------------------- */

/* Replaces constructing ConvexMpc + the OsqpEigen::Solver member
 * (A1RobotControl.cpp:447, A1RobotControl.h:67).  `device` is a CUDA ordinal. */
int mpc_engine_create(const MpcConfig *cfg, int32_t device, MpcEngine **out);
void mpc_engine_destroy(MpcEngine *e);
/* Text of the last error on this engine (or of the last create failure when
 * e == NULL).  Never NULL. */
const char *mpc_last_error(const MpcEngine *e);
/* Make all engine work run on a caller-owned cudaStream_t (e.g. torch's current
 * stream) instead of the engine's own stream. */
int mpc_set_stream(MpcEngine *e, void *cuda_stream);
int mpc_synchronize(MpcEngine *e);
/* Measures this GPU's FP64 FMA issue rate (the roofline denominator of the solvers: B200's tensor
 * cores have no f64 mode worth the name and the path is f64 by its parity bar): a register-only DFMA
 * kernel on every SM for about `ms_target` milliseconds.  Returns TFLOP/s (2 flops per FMA).  bench.py
 * records it next to the clock sample taken while it ran instead of quoting a constant. */
int mpc_measure_fp64_peak(int32_t device, double ms_target, double *tflops);
/* Number of kernels this engine has launched since creation. */
int64_t mpc_kernel_launches(const MpcEngine *e);
/* Developer aid: per-phase SM cycle counters of admm_solve_kernel, summed over CTAs since the
 * last call (out6[0..5]: load+scaling, factorisations, ADMM iterations, residual checks, output,
 * problems; out6[6..21]: fine probes inside the sweep step and the iteration; the buffer must
 * hold 22 values).  enable != 0 switches the counters on for later launches, 0 off. */
int mpc_debug_phase_cycles(MpcEngine *e, int32_t enable, int64_t *out6);

/* ---- batched MPC GRF solve (compute_grf, MPC branch) ---------------------- */

/* K0 loader: host AoS records -> device (A1RobotControl.cpp:452-456 packing
 * happens on the device). */
int mpc_load_states(MpcEngine *e, const MpcStateIn *host, int32_t n);
/* Same, but the records already live in device memory (no copy is made; the
 * buffer must stay valid until results are read). */
int mpc_set_states_device(MpcEngine *e, const MpcStateIn *dev, int32_t n);
/* K1+K2: ConvexMpc::calculate_A_mat_c .. calculate_qp_mats for every loaded
 * state (ConvexMpc.cpp:110-245; A1RobotControl.cpp:472-518). */
int mpc_build_qp(MpcEngine *e);
int mpc_build_qp_async(MpcEngine *e);
/* Parity/debug read-back of one problem: P (n x n row-major), q (n), l, u (m),
 * with n = 12 H, m = 20 H.  Any pointer may be NULL. */
int mpc_get_qp(MpcEngine *e, int32_t idx, float *P, float *q, float *l, float *u);
/* K3+K4+K5: OSQP-equivalent ADMM on every built problem, then the first-step
 * rotation to the body frame (A1RobotControl.cpp:522-561). */
int mpc_solve(MpcEngine *e);
int mpc_solve_async(MpcEngine *e);
/* K5 writer read-back: n results to host. */
int mpc_get_results(MpcEngine *e, MpcResult *host);
/* Device pointer of the result array (n x MpcResult), valid until the next
 * load/solve. */
int mpc_results_device(MpcEngine *e, const MpcResult **dev);
/* Full primal solution of one problem (n floats, world frame, unscaled). */
int mpc_get_solution(MpcEngine *e, int32_t idx, float *x);
/* The whole compute_grf MPC branch for n robots: host records in, host results
 * out (H2D + build + solve + D2H). */
int mpc_compute_grf_batch(MpcEngine *e, const MpcStateIn *host_in, MpcResult *host_out, int32_t n);

/* ---- gait-aware horizon -------------------------------- *
 * The n gait records of the loaded states; required before mpc_build_qp when the engine was
 * created with gait_aware = 1 (MPC_ERR_STATE otherwise), ignored when it was not. */
int mpc_set_gait_inputs(MpcEngine *e, const MpcGaitIn *host, int32_t n);
int mpc_generate_gait_inputs(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick, MpcGaitIn *out);

/* ---- torque map fused into the result writer ----------- *
 * compute_joint_torques (A1RobotControl.cpp:289-319): stance leg tau = J^T (-f_grf), swing leg
 * J tau = km .* f_kin (3x3 partial-pivot LU like Eigen's lu().solve), + torques_gravity; NaN
 * components are flagged, not written.  Give the n records once the states are loaded; the next
 * solve (cold, warm, H = 10 or 30, or the stance-balance QP) writes the torques next to the GRF.
 * NULL switches the map off again.  The first-ten-ticks zero-torque rule (:292-295) is host
 * state and lives in the host mirrors (A1RobotControl::compute_joint_torques). */
int mpc_set_torque_inputs(MpcEngine *e, const MpcTorqueIn *host, int32_t n);
int mpc_get_torques(MpcEngine *e, MpcTorqueOut *host);
/* Synthetic torque-map inputs for the robots of mpc_generate_states: Jacobians of the A1 leg at
 * drawn joint angles (own derivation of the hip-thigh-calf chain, rho_opt = 0 as in
 * GazeboA1ROS.cpp:95), PD forces, km and gravity terms of A1CtrlStates.h:122,129. */
int mpc_generate_torque_inputs(uint64_t seed, uint64_t first_index, int32_t n, MpcTorqueIn *out);

/* ---- upstream state preparation on the device ---------- *
 * One call turns n sensor records into the solver's input, on the device, so that the batch
 * never goes back to the host between sensing and solving:
 *   quaternion -> root_rot_mat, root_euler, yaw rotation       (GazeboA1ROS.cpp:262-269,
 *                                                                Utils.cpp:7-33)
 *   leg forward kinematics and Jacobians, foot positions and
 *   velocities in the robot, rotated and world frames          (GazeboA1ROS.cpp:272-288,
 *                                                                legKinematics/A1Kinematics.cpp)
 *   root_ang_vel = R imu_ang_vel                                (GazeboA1ROS.cpp:306)
 *   A1BasicEKF init_state / update_estimation, one persistent
 *   filter per robot slot                                      (A1BasicEKF.cpp:54-164)
 *   terrain plane fit + moving-window pitch, root_euler_d[1]   (A1RobotControl.cpp:335-376, :566-582)
 * Afterwards the engine is in the state mpc_load_states + mpc_set_torque_inputs would leave it
 * in (Jacobians from the kinematics, PD forces from the record, km / gravity from PrepConfig):
 * call mpc_build_qp / mpc_solve(_warm) / mpc_get_results / mpc_get_torques as usual. */
int prep_config_default(PrepConfig *cfg);
int mpc_prepare_states(MpcEngine *e, const PrepConfig *cfg, const RobotSensorIn *host, int32_t n);
/* Read-back of the prepared records (parity tests, logging); any pointer may be NULL. */
int mpc_get_prepared(MpcEngine *e, MpcStateIn *states, MpcTorqueIn *torque_in, RobotPrepOut *extras);
/* Forget every robot slot's estimator and terrain filter (next tick initialises them). */
int mpc_prepare_reset(MpcEngine *e);
/* Synthetic sensor records for the robots of mpc_generate_stream_states at `tick`. */
int mpc_generate_sensors(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick, RobotSensorIn *out);
/* The A1 leg chain on the host (for tests and host-side users): p[3], J[9] row-major. */
int a1_leg_fk_jac(const double rho_fix[5], const double q[3], double *p, double *J);

/* ---- warm-started streaming: the solver the controller keeps alive --------- *
 * A1RobotControl.h:67 holds ONE OsqpEigen::Solver for the controller's life;
 * A1RobotControl.cpp:522-531 initialises it on the first MPC tick (initSolver,
 * setWarmStart(true)) and :532-538 only updates Hessian, gradient and bounds
 * afterwards, so every later solve starts from the previous tick's iterates
 * and adapted rho.  Problem slot i of the engine is robot i's solver:
 *   - a fresh slot behaves exactly like mpc_solve (cold start);
 *   - a live slot follows OSQP's update semantics: osqp_update_P re-runs the
 *     Ruiz scaling on the new Hessian with the previous gradient still in place,
 *     then gradient and bounds are replaced and rho_vec re-typed; x, z, y stay
 *     in the old scaled coordinates and rho keeps its adapted value.
 * Horizon 10 only (MPC_ERR_UNSUPPORTED otherwise). */
int mpc_solve_warm(MpcEngine *e);
int mpc_solve_warm_async(MpcEngine *e);
/* Forget every live solver (the next warm solve of each slot is an initSolver). */
int mpc_stream_reset(MpcEngine *e);
/* Forget the live solvers of the k robot slots idx[0..k) only (a robot was re-spawned, its state
 * estimate jumped, ...).  The kernels contain faults per slot on their own: a solve that ends with
 * non-finite iterates (one bad sensor record) or an internal error leaves ITS slot dead, so that
 * robot's next tick is an initSolver instead of a warm start from NaN; other robots are untouched. */
int mpc_stream_reset_slots(MpcEngine *e, const int32_t *idx, int32_t k);
/* Replace the model constants of a live engine -- dt, mu, fz bounds, mass, inertia, q / r weights and
 * the solver settings -- WITHOUT touching the live solvers.  The reference re-reads these
 * A1CtrlStates fields on every compute_grf call into a fresh ConvexMpc (A1RobotControl.cpp:447) while
 * its OsqpEigen member solver lives on (A1RobotControl.h:67): a weight change is just another Hessian
 * update.  horizon, the extension flags and structured_solver must match the engine's (MPC_ERR_INVALID). */
int mpc_engine_update_model(MpcEngine *e, const MpcConfig *cfg);
/* One control tick for n robots: load + build + warm solve + results. */
int mpc_stream_step(MpcEngine *e, const MpcStateIn *host_in, MpcResult *host_out, int32_t n);

/* ---- one box, several GPUs: the fleet (SURVEY.md 8b "device list", 8e) -------- *
 * Every QP is independent, so the batch shards with no exchange during build or solve: shard i of G
 * gets the contiguous range [ceil(n i / G), ceil(n (i + 1) / G)) of the caller's arrays.  A fleet owns
 * one engine, one stream and one pinned staging pair per entry of `devices` (a device may be listed
 * more than once: two shards share that GPU).  One call enqueues, for every shard, host-to-device copy,
 * build + solve and the device-to-host copy of the results, then waits for all of them: the results
 * land in ONE caller-owned host array (the "final gather" is these per-GPU copies; no NCCL, no
 * collective).  Caller buffers that are already page-locked are used in place; pageable ones go through
 * the fleet's pinned staging so that the shards' copies overlap.  No CPU fallback: MPC_ERR_NO_DEVICE. */
typedef struct MpcFleet MpcFleet;
int mpc_fleet_create(const MpcConfig *cfg, const int32_t *devices, int32_t ndev, MpcFleet **out);
void mpc_fleet_destroy(MpcFleet *f);
/* Text of the last error on this fleet (of the last create failure when f == NULL).  Never NULL. */
const char *mpc_fleet_last_error(const MpcFleet *f);
int32_t mpc_fleet_size(const MpcFleet *f);
/* Host-only helper: the range of shard i (also what sharding.py uses for torch.distributed ranks). */
int mpc_fleet_shard_range(int32_t n, int32_t ndev, int32_t i, int32_t *begin, int32_t *end);
/* The whole compute_grf MPC branch for n robots over all GPUs of the fleet: host records in, host
 * results out (mpc_compute_grf_batch, sharded). */
int mpc_fleet_compute_grf_batch(MpcFleet *f, const MpcStateIn *host_in, MpcResult *host_out, int32_t n);
/* One warm-started control tick for n robots (mpc_stream_step, sharded; robot i keeps its shard as long
 * as n does not change). */
int mpc_fleet_stream_step(MpcFleet *f, const MpcStateIn *host_in, MpcResult *host_out, int32_t n);
int mpc_fleet_stream_reset(MpcFleet *f);
/* Kernels launched by all engines of the fleet since creation. */
int64_t mpc_fleet_kernel_launches(const MpcFleet *f);

/* ---- ConvexMpc surface, one problem (ConvexMpc.h:22-35) -------------------- */

/* calculate_qp_mats for ONE problem from caller-written model matrices, the way
 * the reference's callers use the public members: A_mat_d (13x13 row-major),
 * B_mat_d_list (13H x 12 row-major, A1RobotControl.cpp:513), mpc_states (13),
 * mpc_states_d (13H), contacts[4].  Outputs hessian (n x n), gradient (n),
 * lb, ub (m).  Runs the general dense build kernel on the device. */
int mpc_qp_mats_from_model(MpcEngine *e, const double *A_mat_d, const double *B_mat_d_list,
                           const double *mpc_states, const double *mpc_states_d,
                           const int32_t *contacts, double *hessian, double *gradient,
                           double *lb, double *ub);
/* OsqpEigen initSolver+solve+getSolution for ONE dense problem with the MPC
 * friction-pyramid constraint matrix (ConvexMpc.cpp:46-58): cold start. */
int mpc_solve_qp(MpcEngine *e, const double *hessian, const double *gradient, const double *lb,
                 const double *ub, double *solution, int32_t *status, int32_t *iters);

/* ---- stance-balance QP (compute_grf, QP branch) --------------------------- */

int balance_engine_create(const BalanceConfig *cfg, int32_t device, MpcEngine **out);
/* A1RobotControl.cpp:377-444 for n robots: host in, host out. */
int balance_qp_solve(MpcEngine *e, const BalanceStateIn *host_in, MpcResult *host_out, int32_t n);
int balance_load_states(MpcEngine *e, const BalanceStateIn *host, int32_t n);
int balance_solve(MpcEngine *e);
int balance_get_qp(MpcEngine *e, int32_t idx, float *P, float *q, float *l, float *u);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
