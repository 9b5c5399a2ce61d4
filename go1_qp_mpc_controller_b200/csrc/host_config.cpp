// Host-side defaults and the synthetic state generator (no device code).
//
// Defaults restate the reference's configuration sources:
//   config/gazebo_a1_mpc.yaml:6-13,40-72   (mass, inertia, q/r weights; primary)
//   config/hardware_a1_mpc.yaml            (well-conditioned weights; secondary)
//   ConvexMpc.cpp:8,223-224                (mu = 0.3, fz in [0,180])
//   A1RobotControl.cpp:11-15,462           (stance-QP constants, mpc_dt)
//   config/gazebo_a1_qp.yaml:54-68         (stance-QP PD gains)
// Solver settings: OSQP 0.6.x library defaults (the reference only switches
// verbosity and warm start, A1RobotControl.cpp:523-524); the benchmark set
// tightens eps to 1e-5 and pins the adaptive-rho interval.
#include <cmath>
#include <cstring>

#include "../../include/mpc_b200.h"

extern "C" {

int mpc_settings_osqp_default(MpcSolverSettings* s) {
  if (!s) return MPC_ERR_INVALID;
  s->rho = 0.1;
  s->sigma = 1e-6;
  s->alpha = 1.6;
  s->eps_abs = 1e-3;
  s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4;
  s->eps_dual_inf = 1e-4;
  s->max_iter = 4000;
  s->check_termination = 25;
  s->scaling = 10;
  s->adaptive_rho = 1;
  // Library default 0 = "choose from setup wall-clock time": not reproducible.
  s->adaptive_rho_interval = 50;
  s->adaptive_rho_tolerance = 5.0;
  return MPC_OK;
}

int mpc_settings_benchmark(MpcSolverSettings* s) {
  int rc = mpc_settings_osqp_default(s);
  if (rc) return rc;
  s->eps_abs = 1e-5;
  s->eps_rel = 1e-5;
  return MPC_OK;
}

static void fill_common(MpcConfig* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->horizon = MPC_HORIZON_DEFAULT;
  cfg->dt = 0.0025;
  cfg->mu = 0.3;
  cfg->fz_min = 0.0;
  cfg->fz_max = 180.0;
  cfg->inertia[0] = 0.0168352186;
  cfg->inertia[4] = 0.0656071082;
  cfg->inertia[8] = 0.0742720659;
  mpc_settings_benchmark(&cfg->osqp);
}

int mpc_config_default(MpcConfig* cfg) {
  if (!cfg) return MPC_ERR_INVALID;
  fill_common(cfg);
  cfg->mass = 12.0;
  const double q[13] = {20.0, 10.0, 1.0, 0.0, 0.0, 420.0, 0.05, 0.05, 0.05, 30.0, 30.0, 10.0, 0.0};
  for (int i = 0; i < 13; ++i) cfg->q_weights[i] = q[i];
  for (int i = 0; i < 12; ++i) cfg->r_weights[i] = 1e-7;
  return MPC_OK;
}

int mpc_config_hardware(MpcConfig* cfg) {
  if (!cfg) return MPC_ERR_INVALID;
  fill_common(cfg);
  cfg->mass = 13.5;
  const double q[13] = {150.0, 150.0, 50.0, 0.0, 0.0, 80.0, 0.2, 0.2, 0.2, 0.3, 0.3, 0.3, 0.0};
  for (int i = 0; i < 13; ++i) cfg->q_weights[i] = q[i];
  for (int i = 0; i < 12; ++i) cfg->r_weights[i] = (i % 3 == 2) ? 0.001 : 0.01;
  return MPC_OK;
}

int balance_config_default(BalanceConfig* cfg) {
  if (!cfg) return MPC_ERR_INVALID;
  std::memset(cfg, 0, sizeof(*cfg));
  const double Q[6] = {1.0, 1.0, 1.0, 400.0, 400.0, 100.0};
  for (int i = 0; i < 6; ++i) cfg->Q[i] = Q[i];
  cfg->R = 1e-3;
  cfg->mu = 0.7;
  cfg->F_min = 0.0;
  cfg->F_max = 180.0;
  cfg->mass = 12.0;
  const double kpl[3] = {100.0, 100.0, 300.0}, kdl[3] = {70.0, 70.0, 120.0};
  const double kpa[3] = {150.0, 150.0, 1.0}, kda[3] = {4.5, 4.5, 30.0};
  for (int i = 0; i < 3; ++i) {
    cfg->kp_linear[i] = kpl[i];
    cfg->kd_linear[i] = kdl[i];
    cfg->kp_angular[i] = kpa[i];
    cfg->kd_angular[i] = kda[i];
  }
  mpc_settings_benchmark(&cfg->osqp);
  return MPC_OK;
}

// ---------------------------------------------------------------------------
// Synthetic robot states.  Counter-based: state i of stream
// `seed` depends only on (seed, i), so any shard of any batch can be produced
// independently on any rank.  splitmix64; u = (x >> 11) * 2^-53.
// ---------------------------------------------------------------------------
namespace {

struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed, uint64_t index)
      : s(seed ^ ((index + 1) * 0x9E3779B97F4A7C15ULL)) {}
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double u01() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uni(double lo, double hi) { return lo + (hi - lo) * u01(); }
};

// R = Rz(yaw) * Ry(pitch) * Rx(roll): the ZYX convention of Utils::quat_to_euler
// (utils/Utils.cpp:7-33).
void euler_to_rot(double roll, double pitch, double yaw, double* R) {
  const double cr = std::cos(roll), sr = std::sin(roll);
  const double cp = std::cos(pitch), sp = std::sin(pitch);
  const double cy = std::cos(yaw), sy = std::sin(yaw);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

// default footholds, config/gazebo_a1_mpc.yaml:17-31 (FL, FR, RL, RR)
const double kDefaultFoot[12] = {0.17, 0.15, -0.35, 0.17, -0.15, -0.35,
                                 -0.17, 0.15, -0.35, -0.17, -0.15, -0.35};

struct CommonDraw {
  double euler[3], pos[3], w[3], v[3], R[9], foot_abs[12], rel[12];
};

void place_feet(CommonDraw& d) {
  euler_to_rot(d.euler[0], d.euler[1], d.euler[2], d.R);
  // foot_pos_abs = R * foot_pos_rel (GazeboA1ROS.cpp:283)
  for (int leg = 0; leg < 4; ++leg)
    for (int r = 0; r < 3; ++r)
      d.foot_abs[3 * leg + r] = d.R[3 * r] * d.rel[3 * leg] + d.R[3 * r + 1] * d.rel[3 * leg + 1] +
                                d.R[3 * r + 2] * d.rel[3 * leg + 2];
}

void draw_common(SplitMix& g, CommonDraw& d) {
  d.euler[0] = g.uni(-0.2, 0.2);
  d.euler[1] = g.uni(-0.2, 0.2);
  d.euler[2] = g.uni(-3.14159265358979323846, 3.14159265358979323846);
  d.pos[0] = g.uni(-1.0, 1.0);
  d.pos[1] = g.uni(-1.0, 1.0);
  d.pos[2] = g.uni(0.22, 0.32);
  for (int i = 0; i < 3; ++i) d.w[i] = g.uni(-1.0, 1.0);
  d.v[0] = g.uni(-0.6, 0.6);
  d.v[1] = g.uni(-0.3, 0.3);
  d.v[2] = g.uni(-0.2, 0.2);
  for (int i = 0; i < 12; ++i) d.rel[i] = kDefaultFoot[i] + g.uni(-0.05, 0.05);
  place_feet(d);
}

}  // namespace

// State `index` of stream `seed`, `tick` control periods (2.5 ms, the main loop period of
// GazeboA1ROS.cpp / HardwareA1ROS.cpp) after its draw: velocities relax towards the commanded
// ones with a 0.1 s time constant, the pose integrates them, trot pairs swap every 48 ticks.
// tick = 0 is exactly the draw.
static void make_mpc_state(uint64_t seed, uint64_t index, int64_t tick, MpcStateIn& s) {
  SplitMix g(seed, index);
  CommonDraw d;
  draw_common(g, d);
  std::memset(&s, 0, sizeof(s));
  s.euler_d[0] = 0.0f;
  s.euler_d[1] = float(g.uni(-0.1, 0.1));
  s.euler_d[2] = 0.0f;
  s.pos_d_z = float(g.uni(0.25, 0.32));        // JOY_CMD_BODY_HEIGHT_MAX, A1Params.h:16
  s.lin_vel_d[0] = float(g.uni(-0.6, 0.6));    // JOY_CMD_VELX_MAX, A1Params.h:19
  s.lin_vel_d[1] = float(g.uni(-0.3, 0.3));    // JOY_CMD_VELY_MAX, A1Params.h:20
  s.lin_vel_d[2] = 0.0f;
  s.ang_vel_d[0] = 0.0f;
  s.ang_vel_d[1] = 0.0f;
  s.ang_vel_d[2] = float(g.uni(-0.8, 0.8));    // JOY_CMD_YAW_MAX, A1Params.h:21
  // trot pairs {FL,RR} / {FR,RL} 45 % each, four-stance 10 %
  const double c = g.u01();
  bool a = c < 0.45, b = c >= 0.45 && c < 0.90;
  if (tick > 0) {
    const double tau = 0.0025 * double(tick);
    const double relax = 1.0 - std::exp(-tau / 0.1);
    for (int i = 0; i < 3; ++i) {
      const double v0 = d.v[i], w0 = d.w[i];
      const double v1 = v0 + (double(s.lin_vel_d[i]) - v0) * relax;
      const double w1 = w0 + (double(s.ang_vel_d[i]) - w0) * relax;
      d.pos[i] += 0.5 * (v0 + v1) * tau;
      d.euler[i] += 0.5 * (w0 + w1) * tau * (i == 2 ? 1.0 : 0.25);
      d.v[i] = v1;
      d.w[i] = w1;
    }
    place_feet(d);
    if (((tick / 48) & 1) && (a || b)) { a = !a; b = !b; }
  }
  for (int i = 0; i < 3; ++i) {
    s.euler[i] = float(d.euler[i]);
    s.pos[i] = float(d.pos[i]);
    s.ang_vel[i] = float(d.w[i]);
    s.lin_vel[i] = float(d.v[i]);
  }
  for (int i = 0; i < 9; ++i) s.rot_mat[i] = float(d.R[i]);
  for (int i = 0; i < 12; ++i) s.foot_pos_abs[i] = float(d.foot_abs[i]);
  s.contacts[0] = (a || (!a && !b)) ? 1.0f : 0.0f;
  s.contacts[3] = s.contacts[0];
  s.contacts[1] = (b || (!a && !b)) ? 1.0f : 0.0f;
  s.contacts[2] = s.contacts[1];
}

int mpc_generate_states(uint64_t seed, uint64_t first_index, int32_t n, MpcStateIn* out) {
  if (!out || n < 0) return MPC_ERR_INVALID;
  for (int32_t k = 0; k < n; ++k) make_mpc_state(seed, first_index + uint64_t(k), 0, out[k]);
  return MPC_OK;
}

int mpc_generate_stream_states(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick,
                               MpcStateIn* out) {
  if (!out || n < 0 || tick < 0) return MPC_ERR_INVALID;
  for (int32_t k = 0; k < n; ++k) make_mpc_state(seed, first_index + uint64_t(k), tick, out[k]);
  return MPC_OK;
}

// ---------------------------------------------------------------------------
// A1 leg kinematics, own derivation of the hip(x)-thigh(y)-calf(y) chain the reference's
// generated code evaluates (legKinematics/A1Kinematics.cpp fk / jac with rho_opt = 0,
// rho_fix = (ox, oy, d, lt, lc), GazeboA1ROS.cpp:76-97):
//   p = (ox, oy, 0) + Rx(q0) [ (0, d, 0) + Ry(q1) ( (0,0,-lt) + Ry(q2) (0,0,-lc) ) ]
// J = dp/dq, row-major.
// ---------------------------------------------------------------------------
int a1_leg_fk_jac(const double rho_fix[5], const double q[3], double* p, double* J) {
  if (!rho_fix || !q) return MPC_ERR_INVALID;
  const double ox = rho_fix[0], oy = rho_fix[1], d = rho_fix[2], lt = rho_fix[3], lc = rho_fix[4];
  const double c0 = std::cos(q[0]), s0 = std::sin(q[0]);
  const double c1 = std::cos(q[1]), s1 = std::sin(q[1]);
  const double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
  const double L = lt * c1 + lc * c12;    // leg extension below the hip axis
  const double X = -lt * s1 - lc * s12;   // forward reach = dL/dq1
  if (p) {
    p[0] = ox + X;
    p[1] = oy + d * c0 + L * s0;
    p[2] = d * s0 - L * c0;
  }
  if (J) {
    J[0] = 0.0;              J[1] = -L;       J[2] = -lc * c12;
    J[3] = -d * s0 + L * c0; J[4] = s0 * X;   J[5] = -s0 * lc * s12;
    J[6] = d * c0 + L * s0;  J[7] = -c0 * X;  J[8] = c0 * lc * s12;
  }
  return MPC_OK;
}

int prep_config_default(PrepConfig* cfg) {
  if (!cfg) return MPC_ERR_INVALID;
  std::memset(cfg, 0, sizeof(*cfg));
  for (int leg = 0; leg < 4; ++leg) {
    double* r = cfg->rho_fix + 5 * leg;                 // GazeboA1ROS.cpp:76-93
    r[0] = (leg < 2) ? 0.1881 : -0.1881;
    r[1] = (leg % 2 == 0) ? 0.04675 : -0.04675;
    r[2] = (leg % 2 == 0) ? 0.08 : -0.08;
    r[3] = 0.213;
    r[4] = 0.213;
    cfg->torques_gravity[3 * leg] = (leg % 2 == 0) ? 0.80 : -0.80;  // A1CtrlStates.h:129
  }
  cfg->km_foot[0] = cfg->km_foot[1] = cfg->km_foot[2] = 0.1;        // A1CtrlStates.h:122
  cfg->use_estimator = 1;
  cfg->assume_flat_ground = 1;
  cfg->use_terrain_adapt = 1;
  return MPC_OK;
}

int mpc_generate_torque_inputs(uint64_t seed, uint64_t first_index, int32_t n, MpcTorqueIn* out) {
  if (!out || n < 0) return MPC_ERR_INVALID;
  for (int32_t k = 0; k < n; ++k) {
    SplitMix g(seed ^ 0x746f727175653a31ULL, first_index + uint64_t(k));  // a stream of its own
    MpcTorqueIn& t = out[k];
    std::memset(&t, 0, sizeof(t));
    for (int leg = 0; leg < 4; ++leg) {
      const double q[3] = {g.uni(-0.3, 0.3), g.uni(0.5, 1.1), g.uni(-2.0, -1.2)};
      double p[3], J[9];
      PrepConfig pc;
      prep_config_default(&pc);
      a1_leg_fk_jac(pc.rho_fix + 5 * leg, q, p, J);
      for (int i = 0; i < 9; ++i) t.j_foot[9 * leg + i] = float(J[i]);
      t.foot_forces_kin[3 * leg + 0] = float(g.uni(-40.0, 40.0));
      t.foot_forces_kin[3 * leg + 1] = float(g.uni(-40.0, 40.0));
      t.foot_forces_kin[3 * leg + 2] = float(g.uni(-60.0, 60.0));
      t.torques_gravity[3 * leg] = (leg % 2 == 0) ? 0.80f : -0.80f;  // A1CtrlStates.h:129
    }
    t.km_foot[0] = t.km_foot[1] = t.km_foot[2] = 0.1f;               // A1CtrlStates.h:122
  }
  return MPC_OK;
}

// Sensor records consistent with make_mpc_state(seed, index, tick): the same pose, velocities,
// commands and contacts, seen through quaternion / IMU / joint encoders / foot force sensors.
int mpc_generate_sensors(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick, RobotSensorIn* out) {
  if (!out || n < 0 || tick < 0) return MPC_ERR_INVALID;
  PrepConfig pc;
  prep_config_default(&pc);
  for (int32_t k = 0; k < n; ++k) {
    const uint64_t index = first_index + uint64_t(k);
    MpcStateIn st;
    make_mpc_state(seed, index, tick, st);
    SplitMix g(seed ^ 0x73656e736f72733aULL, index);  // per-robot constants of the sensor stream
    RobotSensorIn& s = out[k];
    std::memset(&s, 0, sizeof(s));
    const double tau = 0.0025 * double(tick);
    // quaternion of the ZYX euler angles
    const double cr = std::cos(0.5 * st.euler[0]), sr = std::sin(0.5 * st.euler[0]);
    const double cp = std::cos(0.5 * st.euler[1]), sp = std::sin(0.5 * st.euler[1]);
    const double cy = std::cos(0.5 * st.euler[2]), sy = std::sin(0.5 * st.euler[2]);
    s.root_quat[0] = float(cr * cp * cy + sr * sp * sy);
    s.root_quat[1] = float(sr * cp * cy - cr * sp * sy);
    s.root_quat[2] = float(cr * sp * cy + sr * cp * sy);
    s.root_quat[3] = float(cr * cp * sy - sr * sp * cy);
    const float* R = st.rot_mat;
    // IMU: body-frame angular velocity and specific force (gravity + a slow sway)
    const double acc_w[3] = {0.4 * std::sin(7.0 * tau + g.uni(0, 6.28)), 0.4 * std::cos(5.0 * tau + g.uni(0, 6.28)),
                             9.81 + 0.3 * std::sin(9.0 * tau + g.uni(0, 6.28))};
    for (int c = 0; c < 3; ++c) {
      s.imu_ang_vel[c] = float(R[c] * st.ang_vel[0] + R[3 + c] * st.ang_vel[1] + R[6 + c] * st.ang_vel[2]);
      s.imu_acc[c] = float(R[c] * acc_w[0] + R[3 + c] * acc_w[1] + R[6 + c] * acc_w[2]);
    }
    const double slope_x = g.uni(-0.25, 0.25), slope_y = g.uni(-0.1, 0.1);  // terrain under the feet
    for (int leg = 0; leg < 4; ++leg) {
      const double ph = g.uni(0, 6.28);
      const double q[3] = {g.uni(-0.25, 0.25) + 0.05 * std::sin(6.0 * tau + ph), g.uni(0.6, 1.0) + 0.1 * std::sin(8.0 * tau + ph),
                           g.uni(-1.9, -1.3) + 0.1 * std::cos(8.0 * tau + ph)};
      const double qd[3] = {0.3 * std::cos(6.0 * tau + ph), 0.8 * std::cos(8.0 * tau + ph), -0.8 * std::sin(8.0 * tau + ph)};
      for (int c = 0; c < 3; ++c) { s.joint_pos[3 * leg + c] = float(q[c]); s.joint_vel[3 * leg + c] = float(qd[c]); }
      const bool contact = st.contacts[leg] != 0.0f;
      s.foot_force[leg] = contact ? float(g.uni(60.0, 160.0)) : float(g.uni(0.0, 30.0));
      double p[3];
      a1_leg_fk_jac(pc.rho_fix + 5 * leg, q, p, nullptr);
      s.foot_pos_recent_contact[3 * leg] = float(p[0]);
      s.foot_pos_recent_contact[3 * leg + 1] = float(p[1]);
      s.foot_pos_recent_contact[3 * leg + 2] = float(-0.30 + slope_x * p[0] + slope_y * p[1] + g.uni(-0.005, 0.005));
      s.foot_forces_kin[3 * leg] = float(g.uni(-40.0, 40.0));
      s.foot_forces_kin[3 * leg + 1] = float(g.uni(-40.0, 40.0));
      s.foot_forces_kin[3 * leg + 2] = float(g.uni(-60.0, 60.0));
      s.contacts[leg] = st.contacts[leg];
    }
    for (int c = 0; c < 3; ++c) {
      s.root_pos[c] = st.pos[c];
      s.root_lin_vel[c] = st.lin_vel[c];
      s.root_euler_d[c] = st.euler_d[c];
      s.root_lin_vel_d[c] = st.lin_vel_d[c];
      s.root_ang_vel_d[c] = st.ang_vel_d[c];
    }
    s.root_pos_d_z = st.pos_d_z;
    s.movement_mode = (g.u01() < 0.9) ? 1.0f : 0.0f;
    s.dt = 0.0025f;
  }
  return MPC_OK;
}

// Gait scheduler records consistent with make_mpc_state's trot pattern: counters advance by
// gait_counter_speed = 2 per tick through a 240-count gait with 120 counts of stance
// (A1CtrlStates.h:24-25,103); the phase is chosen so that "counter <= counter_per_swing" agrees
// with the state's contact flags at this tick and swaps together with them every 48 ticks...
// except that a real gait swaps every 60 ticks: records near a swap exercise contact changes
// INSIDE the horizon, which is what the gait-aware bounds are for.
int mpc_generate_gait_inputs(uint64_t seed, uint64_t first_index, int32_t n, int64_t tick, MpcGaitIn* out) {
  if (!out || n < 0 || tick < 0) return MPC_ERR_INVALID;
  for (int32_t k = 0; k < n; ++k) {
    const uint64_t index = first_index + uint64_t(k);
    MpcStateIn st;
    make_mpc_state(seed, index, tick, st);
    SplitMix g(seed ^ 0x676169743a303031ULL, index);
    MpcGaitIn& r = out[k];
    std::memset(&r, 0, sizeof(r));
    r.counter_per_gait = 240.0f;
    r.counter_per_swing = 120.0f;
    r.ticks_per_step = 1.0f;
    const double into = g.uni(0.0, 118.0);  // how far the current phase (stance or swing) has progressed
    for (int leg = 0; leg < 4; ++leg) {
      r.gait_counter_speed[leg] = 2.0f;
      const bool contact = st.contacts[leg] != 0.0f;
      const bool all_four = st.contacts[0] != 0.0f && st.contacts[1] != 0.0f;
      // stance occupies counts [0, 120], swing (120, 240); a standing robot stays deep in stance
      double c = contact ? std::floor(into) : 121.0 + std::floor(into);
      if (all_four) c = 10.0;
      r.gait_counter[leg] = float(c);
    }
  }
  return MPC_OK;
}

int balance_generate_states(uint64_t seed, uint64_t first_index, int32_t n, BalanceStateIn* out) {
  if (!out || n < 0) return MPC_ERR_INVALID;
  for (int32_t k = 0; k < n; ++k) {
    SplitMix g(seed, first_index + uint64_t(k));
    CommonDraw d;
    draw_common(g, d);
    BalanceStateIn& s = out[k];
    std::memset(&s, 0, sizeof(s));
    for (int i = 0; i < 3; ++i) {
      s.euler[i] = float(d.euler[i]);
      s.pos[i] = float(d.pos[i]);
      s.ang_vel[i] = float(d.w[i]);
      s.lin_vel[i] = float(d.v[i]);
    }
    s.euler_d[0] = 0.0f;
    s.euler_d[1] = float(g.uni(-0.1, 0.1));
    // desired yaw follows the current yaw plus a small command offset
    s.euler_d[2] = float(d.euler[2] + g.uni(-0.2, 0.2));
    s.pos_d[0] = float(d.pos[0]);
    s.pos_d[1] = float(d.pos[1]);
    s.pos_d[2] = float(g.uni(0.25, 0.32));
    s.lin_vel_d[0] = float(g.uni(-0.6, 0.6));
    s.lin_vel_d[1] = float(g.uni(-0.3, 0.3));
    s.lin_vel_d[2] = 0.0f;
    s.ang_vel_d[2] = float(g.uni(-0.8, 0.8));
    for (int i = 0; i < 9; ++i) s.rot_mat[i] = float(d.R[i]);
    double Rz[9];
    euler_to_rot(0.0, 0.0, d.euler[2], Rz);
    for (int i = 0; i < 9; ++i) s.rot_mat_z[i] = float(Rz[i]);
    for (int i = 0; i < 12; ++i) s.foot_pos_abs[i] = float(d.foot_abs[i]);
    // four-stance 80 %, trot pairs 10 % each
    const double c = g.u01();
    const bool all4 = c < 0.80, a = c >= 0.80 && c < 0.90;
    s.contacts[0] = (all4 || a) ? 1.0f : 0.0f;
    s.contacts[3] = s.contacts[0];
    s.contacts[1] = (all4 || !a) ? 1.0f : 0.0f;
    s.contacts[2] = s.contacts[1];
  }
  return MPC_OK;
}

}  // extern "C"
