// convex_mpc_b200.hpp -- header-only C++ mirror of the reference's interface for the hot
// path, implemented over the C ABI of mpc_b200.h.  Eigen does not exist in the build image,
// so matrices are plain row-major std::array / std::vector; field and method names follow
//   class ConvexMpc                 src/a1_cpp/src/ConvexMpc.h:22-92
//   A1CtrlStates (subset)           src/a1_cpp/src/A1CtrlStates.h:330-447
//   A1RobotControl::compute_grf     src/a1_cpp/src/A1RobotControl.h:44
// A maintainer with Eigen swaps the containers for Eigen::Map<> views.
#ifndef CONVEX_MPC_B200_HPP
#define CONVEX_MPC_B200_HPP

#include <array>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpc_b200.h"

namespace mpc_b200 {

constexpr int PLAN_HORIZON = MPC_HORIZON_DEFAULT;
constexpr int NUM_LEG = MPC_NUM_LEG;
constexpr int NUM_DOF = MPC_NUM_DOF;

inline void check(int rc, const MpcEngine* e) {
  if (rc != MPC_OK) throw std::runtime_error(std::string("mpc_b200: ") + mpc_last_error(e));
}

// The A1CtrlStates fields compute_grf reads and writes; 3x3 row-major, foot_pos_* column-per-leg
// like the reference (3 x NUM_LEG, stored row-major).
struct A1CtrlStates {
  int stance_leg_control_type = 1;  // 0: QP, 1: MPC
  int use_terrain_adapt = 1;
  double robot_mass = 15.0;
  std::array<double, 9> a1_trunk_inertia{0.0168352186, 0, 0, 0, 0.0656071082, 0, 0, 0, 0.0742720659};
  std::array<double, 13> q_weights{80, 80, 1, 0, 0, 270, 1, 1, 20, 20, 20, 20, 0};
  std::array<double, 12> r_weights{1e-5, 1e-5, 1e-6, 1e-5, 1e-5, 1e-6, 1e-5, 1e-5, 1e-6, 1e-5, 1e-5, 1e-6};
  std::array<double, 3> root_pos{}, root_euler{}, root_lin_vel{}, root_ang_vel{};
  std::array<double, 3> root_pos_d{}, root_euler_d{}, root_lin_vel_d{}, root_lin_vel_d_world{}, root_ang_vel_d{};
  std::array<double, 9> root_rot_mat{}, root_rot_mat_z{};
  std::array<double, 12> foot_pos_abs{};  // 3 x 4, element (r, leg) at [4 r + leg]
  bool contacts[NUM_LEG] = {false, false, false, false};
  std::array<double, 13> mpc_states{};
  std::vector<double> mpc_states_d = std::vector<double>(13 * PLAN_HORIZON, 0.0);
  std::array<double, 3> kp_linear{1000, 1000, 1000}, kd_linear{200, 70, 120};
  std::array<double, 3> kp_angular{650, 35, 1}, kd_angular{4.5, 4.5, 30};
  // compute_joint_torques inputs / output (A1CtrlStates.h:95,122,129)
  std::array<double, 36> j_foot_blocks{1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 1,
                                       1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 1};  // 4 x (3x3 row-major)
  std::array<double, 12> foot_forces_kin{};  // 3 x 4, element (r, leg) at [4 r + leg]
  std::array<double, 3> km_foot{0.1, 0.1, 0.1};
  std::array<double, 12> torques_gravity{0.80, 0, 0, -0.80, 0, 0, 0.80, 0, 0, -0.80, 0, 0};
  std::array<double, 12> joint_torques{};
  MpcTorqueIn to_torque_record() const {
    MpcTorqueIn t;
    std::memset(&t, 0, sizeof(t));
    for (int i = 0; i < 36; ++i) t.j_foot[i] = (float)j_foot_blocks[i];
    for (int leg = 0; leg < 4; ++leg)
      for (int k = 0; k < 3; ++k) t.foot_forces_kin[3 * leg + k] = (float)foot_forces_kin[4 * k + leg];
    for (int i = 0; i < 3; ++i) t.km_foot[i] = (float)km_foot[i];
    for (int i = 0; i < 12; ++i) t.torques_gravity[i] = (float)torques_gravity[i];
    return t;
  }

  MpcStateIn to_record() const {
    MpcStateIn r;
    std::memset(&r, 0, sizeof(r));
    for (int i = 0; i < 3; ++i) {
      r.euler[i] = (float)root_euler[i]; r.pos[i] = (float)root_pos[i];
      r.ang_vel[i] = (float)root_ang_vel[i]; r.lin_vel[i] = (float)root_lin_vel[i];
      r.euler_d[i] = (float)root_euler_d[i]; r.lin_vel_d[i] = (float)root_lin_vel_d[i];
      r.ang_vel_d[i] = (float)root_ang_vel_d[i];
    }
    r.pos_d_z = (float)root_pos_d[2];
    for (int i = 0; i < 9; ++i) r.rot_mat[i] = (float)root_rot_mat[i];
    for (int leg = 0; leg < 4; ++leg)
      for (int k = 0; k < 3; ++k) r.foot_pos_abs[3 * leg + k] = (float)foot_pos_abs[4 * k + leg];
    for (int i = 0; i < 4; ++i) r.contacts[i] = contacts[i] ? 1.0f : 0.0f;
    return r;
  }
  BalanceStateIn to_balance_record() const {
    BalanceStateIn r;
    std::memset(&r, 0, sizeof(r));
    for (int i = 0; i < 3; ++i) {
      r.euler[i] = (float)root_euler[i]; r.pos[i] = (float)root_pos[i];
      r.ang_vel[i] = (float)root_ang_vel[i]; r.lin_vel[i] = (float)root_lin_vel[i];
      r.euler_d[i] = (float)root_euler_d[i]; r.pos_d[i] = (float)root_pos_d[i];
      r.lin_vel_d[i] = (float)root_lin_vel_d[i]; r.ang_vel_d[i] = (float)root_ang_vel_d[i];
    }
    for (int i = 0; i < 9; ++i) { r.rot_mat[i] = (float)root_rot_mat[i]; r.rot_mat_z[i] = (float)root_rot_mat_z[i]; }
    for (int leg = 0; leg < 4; ++leg)
      for (int k = 0; k < 3; ++k) r.foot_pos_abs[3 * leg + k] = (float)foot_pos_abs[4 * k + leg];
    for (int i = 0; i < 4; ++i) r.contacts[i] = contacts[i] ? 1.0f : 0.0f;
    return r;
  }
};

// ConvexMpc with the reference's five methods and public data members.
class ConvexMpc {
 public:
  ConvexMpc(const std::array<double, 13>& q_weights_, const std::array<double, 12>& r_weights_, int device = 0) {
    mu = 0.3; fz_min = 0.0; fz_max = 0.0;                       // ConvexMpc.cpp:8-10
    // linear_constraints (ConvexMpc.cpp:46-58), dense row-major (20 H) x (12 H): per leg-step the rows
    // fx + mu fz, fx - mu fz, fy + mu fz, fy - mu fz, fz.  The reference's callers read it when they
    // set up their solver (A1RobotControl.cpp:527, test/test_mpc.cpp:143).
    linear_constraints.assign(size_t(20 * PLAN_HORIZON) * 12 * PLAN_HORIZON, 0.0);
    for (int i = 0; i < NUM_LEG * PLAN_HORIZON; ++i) {
      auto lc = [&](int r, int c) -> double& { return linear_constraints[size_t(r) * 12 * PLAN_HORIZON + c]; };
      lc(0 + 5 * i, 0 + 3 * i) = 1; lc(1 + 5 * i, 0 + 3 * i) = 1;
      lc(2 + 5 * i, 1 + 3 * i) = 1; lc(3 + 5 * i, 1 + 3 * i) = 1;
      lc(4 + 5 * i, 2 + 3 * i) = 1;
      lc(0 + 5 * i, 2 + 3 * i) = mu; lc(1 + 5 * i, 2 + 3 * i) = -mu;
      lc(2 + 5 * i, 2 + 3 * i) = mu; lc(3 + 5 * i, 2 + 3 * i) = -mu;
    }
    MpcConfig cfg;
    mpc_config_default(&cfg);
    for (int i = 0; i < 13; ++i) cfg.q_weights[i] = q_weights_[i];
    for (int i = 0; i < 12; ++i) cfg.r_weights[i] = r_weights_[i];
    int rc = mpc_engine_create(&cfg, device, &engine_);
    if (rc != MPC_OK) throw std::runtime_error(std::string("mpc_b200: ") + mpc_last_error(nullptr));
    reset();
  }
  ~ConvexMpc() { mpc_engine_destroy(engine_); }
  ConvexMpc(const ConvexMpc&) = delete;
  ConvexMpc& operator=(const ConvexMpc&) = delete;

  void reset() {                                                 // ConvexMpc.cpp:70-108
    A_mat_c.fill(0); B_mat_c.fill(0); A_mat_d.fill(0); B_mat_d.fill(0);
    B_mat_d_list.assign(13 * PLAN_HORIZON * 12, 0.0);
    hessian.assign(size_t(12 * PLAN_HORIZON) * 12 * PLAN_HORIZON, 0.0);
    gradient.assign(12 * PLAN_HORIZON, 0.0);
    lb.assign(20 * PLAN_HORIZON, 0.0);
    ub.assign(20 * PLAN_HORIZON, 0.0);
  }
  void calculate_A_mat_c(const std::array<double, 3>& root_euler) {  // ConvexMpc.cpp:110-130
    const double cy = std::cos(root_euler[2]), sy = std::sin(root_euler[2]);
    auto A = [&](int r, int c) -> double& { return A_mat_c[r * 13 + c]; };
    A(0, 6) = cy; A(0, 7) = sy; A(0, 8) = 0;
    A(1, 6) = -sy; A(1, 7) = cy; A(1, 8) = 0;
    A(2, 6) = 0; A(2, 7) = 0; A(2, 8) = 1;
    A(3, 9) = 1; A(4, 10) = 1; A(5, 11) = 1;
    A(11, NUM_DOF) = 1;
  }
  // foot_pos: 3 x 4 row-major (column per leg), ConvexMpc.cpp:132-143
  void calculate_B_mat_c(double robot_mass, const std::array<double, 9>& a1_trunk_inertia,
                         const std::array<double, 9>& R, const std::array<double, 12>& foot_pos) {
    double T[9], Iw[9], inv[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += R[3 * i + k] * a1_trunk_inertia[3 * k + j];
        T[3 * i + j] = s;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += T[3 * i + k] * R[3 * j + k];
        Iw[3 * i + j] = s;
      }
    const double c00 = Iw[4] * Iw[8] - Iw[5] * Iw[7], c01 = Iw[5] * Iw[6] - Iw[3] * Iw[8],
                 c02 = Iw[3] * Iw[7] - Iw[4] * Iw[6];
    const double id = 1.0 / (Iw[0] * c00 + Iw[1] * c01 + Iw[2] * c02);
    inv[0] = c00 * id; inv[1] = (Iw[2] * Iw[7] - Iw[1] * Iw[8]) * id; inv[2] = (Iw[1] * Iw[5] - Iw[2] * Iw[4]) * id;
    inv[3] = c01 * id; inv[4] = (Iw[0] * Iw[8] - Iw[2] * Iw[6]) * id; inv[5] = (Iw[2] * Iw[3] - Iw[0] * Iw[5]) * id;
    inv[6] = c02 * id; inv[7] = (Iw[1] * Iw[6] - Iw[0] * Iw[7]) * id; inv[8] = (Iw[0] * Iw[4] - Iw[1] * Iw[3]) * id;
    for (int leg = 0; leg < NUM_LEG; ++leg) {
      const double x = foot_pos[leg], y = foot_pos[4 + leg], z = foot_pos[8 + leg];
      const double sk[9] = {0, -z, y, z, 0, -x, -y, x, 0};       // Utils::skew
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += inv[3 * i + k] * sk[3 * k + j];
          B_mat_c[(6 + i) * 12 + 3 * leg + j] = s;
          B_mat_c[(9 + i) * 12 + 3 * leg + j] = (i == j) ? 1.0 / robot_mass : 0.0;
        }
    }
  }
  void state_space_discretization(double dt) {                   // ConvexMpc.cpp:145-156
    for (int r = 0; r < 13; ++r)
      for (int c = 0; c < 13; ++c) A_mat_d[r * 13 + c] = (r == c ? 1.0 : 0.0) + A_mat_c[r * 13 + c] * dt;
    for (int i = 0; i < 13 * 12; ++i) B_mat_d[i] = B_mat_c[i] * dt;
  }
  // ConvexMpc.cpp:158-245, on the GPU (general dense build kernel)
  void calculate_qp_mats(A1CtrlStates& state) {
    int32_t c[4];
    for (int i = 0; i < 4; ++i) c[i] = state.contacts[i] ? 1 : 0;
    check(mpc_qp_mats_from_model(engine_, A_mat_d.data(), B_mat_d_list.data(), state.mpc_states.data(),
                                 state.mpc_states_d.data(), c, hessian.data(), gradient.data(), lb.data(),
                                 ub.data()), engine_);
    fz_min = 0; fz_max = 180;
  }
  // OsqpEigen::Solver::{initSolver, solve, getSolution} on the members above, cold start
  std::vector<double> solve(int* status = nullptr, int* iters = nullptr) {
    std::vector<double> x(12 * PLAN_HORIZON);
    int32_t st = 0, it = 0;
    check(mpc_solve_qp(engine_, hessian.data(), gradient.data(), lb.data(), ub.data(), x.data(), &st, &it), engine_);
    if (status) *status = st;
    if (iters) *iters = it;
    return x;
  }

  double mu, fz_min, fz_max;
  std::array<double, 13 * 13> A_mat_c, A_mat_d;
  std::array<double, 13 * 12> B_mat_c, B_mat_d;
  std::vector<double> B_mat_d_list;  // (13 H) x 12, written by the caller (A1RobotControl.cpp:513)
  std::vector<double> hessian, gradient, lb, ub;
  std::vector<double> linear_constraints;  // (20 H) x (12 H) row-major, constant (ConvexMpc.h:84)

 private:
  MpcEngine* engine_ = nullptr;
};

// compute_grf for one robot or a batch of robots.
class A1RobotControl {
 public:
  explicit A1RobotControl(int device = 0) : device_(device) {}
  ~A1RobotControl() { mpc_engine_destroy(mpc_); mpc_engine_destroy(qp_); }

  // returns the 3 x 4 body-frame GRF (row-major), like the reference
  std::array<double, 12> compute_grf(A1CtrlStates& state, double dt) {
    MpcResult res;
    if (state.stance_leg_control_type == 1) {
      const double mpc_dt = (use_sim_time == "true") ? dt : 0.0025;          // A1RobotControl.cpp:462-467
      ensure_mpc(state, mpc_dt);
      for (int i = 0; i < 3; ++i) {
        state.mpc_states[i] = state.root_euler[i]; state.mpc_states[3 + i] = state.root_pos[i];
        state.mpc_states[6 + i] = state.root_ang_vel[i]; state.mpc_states[9 + i] = state.root_lin_vel[i];
      }
      state.mpc_states[12] = -9.8;
      MpcStateIn rec = state.to_record();
      MpcTorqueIn tin = state.to_torque_record();
      check(mpc_load_states(mpc_, &rec, 1), mpc_);
      check(mpc_set_torque_inputs(mpc_, &tin, 1), mpc_);
      check(mpc_build_qp_async(mpc_), mpc_);
      // slot 0 of the engine is the member solver of A1RobotControl.h:67: initSolver on the first
      // tick, update* + warm solve() on every later one (A1RobotControl.cpp:522-540)
      check(mpc_solve_warm_async(mpc_), mpc_);
      check(mpc_get_results(mpc_, &res), mpc_);
      check(mpc_get_torques(mpc_, &torques_), mpc_);
    } else {
      ensure_qp(state);
      BalanceStateIn rec = state.to_balance_record();
      MpcTorqueIn tin = state.to_torque_record();
      check(balance_load_states(qp_, &rec, 1), qp_);
      check(mpc_set_torque_inputs(qp_, &tin, 1), qp_);
      check(balance_solve(qp_), qp_);
      check(mpc_get_results(qp_, &res), qp_);
      check(mpc_get_torques(qp_, &torques_), qp_);
    }
    have_torques_ = true;
    last_status = res.status;
    last_iters = res.iters;
    std::array<double, 12> grf;
    for (int leg = 0; leg < 4; ++leg)
      for (int k = 0; k < 3; ++k) grf[4 * k + leg] = res.grf[3 * leg + k];
    return grf;
  }
  // A1RobotControl.cpp:289-319 with the torques the device wrote next to the last GRF: zero for
  // the first nine calls, NaN components keep their previous value
  void compute_joint_torques(A1CtrlStates& state) {
    if (++mpc_init_counter < 10) {
      state.joint_torques.fill(0.0);
      return;
    }
    if (!have_torques_) throw std::runtime_error("compute_joint_torques before compute_grf");
    for (int i = 0; i < 12; ++i)
      if (!((torques_.nan_mask >> i) & 1)) state.joint_torques[i] = torques_.joint_torques[i];
  }
  int mpc_init_counter = 0;
  int last_status = 0, last_iters = 0;  // OSQP status / iteration count of the last compute_grf (the reference drops them)
  // forget the member solver: the next MPC call is an initSolver again
  void reset_solver() { if (mpc_) check(mpc_stream_reset(mpc_), mpc_); }
  // the MPC branch for n robots sharing the engine-wide constants of `cfg`
  void compute_grf_batch(const MpcConfig& cfg, const MpcStateIn* states, MpcResult* out, int n) {
    if (!mpc_) check(mpc_engine_create(&cfg, device_, &mpc_), nullptr);
    check(mpc_compute_grf_batch(mpc_, states, out, n), mpc_);
  }
  std::string use_sim_time = "false";

 private:
  void ensure_mpc(const A1CtrlStates& s, double mpc_dt) {
    MpcConfig cfg;
    mpc_config_default(&cfg);
    cfg.dt = mpc_dt; cfg.mass = s.robot_mass;
    for (int i = 0; i < 9; ++i) cfg.inertia[i] = s.a1_trunk_inertia[i];
    for (int i = 0; i < 13; ++i) cfg.q_weights[i] = s.q_weights[i];
    for (int i = 0; i < 12; ++i) cfg.r_weights[i] = s.r_weights[i];
    if (mpc_ && std::memcmp(&cfg, &cfg_, sizeof(cfg)) == 0) return;
    if (mpc_) {
      // the reference re-reads these fields into a fresh ConvexMpc on every call while its solver lives
      // on (A1RobotControl.cpp:447, A1RobotControl.h:67): new constants, same warm solver
      check(mpc_engine_update_model(mpc_, &cfg), mpc_);
    } else {
      check(mpc_engine_create(&cfg, device_, &mpc_), nullptr);
    }
    cfg_ = cfg;
  }
  void ensure_qp(const A1CtrlStates& s) {
    // gains and mass are read from the state on every tick (A1RobotControl.cpp:380-391)
    BalanceConfig b;
    balance_config_default(&b);
    b.mass = s.robot_mass;
    for (int i = 0; i < 3; ++i) {
      b.kp_linear[i] = s.kp_linear[i]; b.kd_linear[i] = s.kd_linear[i];
      b.kp_angular[i] = s.kp_angular[i]; b.kd_angular[i] = s.kd_angular[i];
    }
    if (qp_ && std::memcmp(&b, &bcfg_, sizeof(b)) == 0) return;
    mpc_engine_destroy(qp_);
    qp_ = nullptr;
    check(balance_engine_create(&b, device_, &qp_), nullptr);
    bcfg_ = b;
  }
  int device_;
  MpcEngine* mpc_ = nullptr;
  MpcEngine* qp_ = nullptr;
  MpcConfig cfg_{};
  BalanceConfig bcfg_{};
  MpcTorqueOut torques_{};
  bool have_torques_ = false;
};

}  // namespace mpc_b200
#endif
