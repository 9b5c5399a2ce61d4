// The reference's stand-alone driver (src/a1_cpp/src/test/test_mpc.cpp:14-159) restated against
// include/convex_mpc_b200.hpp: same state literal, same call sequence, prints the 3x4 GRF.
// Exit code 0 iff the solve reports OSQP_SOLVED and FL/RL carry the known answer
// (SURVEY.md 8c probe: (0, -12.78, 42.61) N at eps 1e-5).
#include <chrono>
#include <cmath>
#include <cstdio>

#include "convex_mpc_b200.hpp"

using namespace mpc_b200;

int main() {
  A1CtrlStates state;
  state.robot_mass = 15;
  state.a1_trunk_inertia = {0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542};
  state.root_euler = {0, 0, 0};
  state.root_rot_mat = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  state.root_pos = {0, 0, 0.15};
  const std::array<double, 12> foot_pos_rel = {0.17, 0.17, -0.17, -0.17, 0.15, -0.15, 0.15, -0.15,
                                               -0.35, -0.35, -0.35, -0.35};
  state.contacts[0] = true; state.contacts[1] = false; state.contacts[2] = true; state.contacts[3] = false;
  const double dt = 0.0025;
  std::array<double, 13> q_weights = {1, 1, 1, 0, 0, 50, 0, 0, 1, 1, 1, 1, 0};
  std::array<double, 12> r_weights;
  r_weights.fill(1e-6);
  try {
    ConvexMpc mpc_solver(q_weights, r_weights);
    mpc_solver.reset();
    for (int i = 0; i < 3; ++i) {
      state.mpc_states[i] = state.root_euler[i]; state.mpc_states[3 + i] = state.root_pos[i];
      state.mpc_states[6 + i] = state.root_ang_vel[i]; state.mpc_states[9 + i] = state.root_lin_vel[i];
    }
    state.mpc_states[12] = -9.8;
    for (int i = 0; i < PLAN_HORIZON; ++i) {
      double* d = &state.mpc_states_d[13 * i];
      for (int k = 0; k < 13; ++k) d[k] = 0.0;
      d[5] = state.root_pos[2];
      d[12] = -9.8;
    }
    mpc_solver.calculate_A_mat_c({0, 0, 0});
    for (int i = 0; i < PLAN_HORIZON; ++i) {
      mpc_solver.calculate_B_mat_c(state.robot_mass, state.a1_trunk_inertia, state.root_rot_mat, foot_pos_rel);
      mpc_solver.state_space_discretization(dt);
      for (int k = 0; k < 13 * 12; ++k) mpc_solver.B_mat_d_list[13 * 12 * i + k] = mpc_solver.B_mat_d[k];
    }
    mpc_solver.calculate_qp_mats(state);
    auto t0 = std::chrono::high_resolution_clock::now();
    int status = 0, iters = 0;
    std::vector<double> sol = mpc_solver.solve(&status, &iters);
    auto t1 = std::chrono::high_resolution_clock::now();
    for (int r = 0; r < 3; ++r) {
      for (int leg = 0; leg < 4; ++leg) std::printf("%10.4f ", sol[3 * leg + r]);
      std::printf("\n");
    }
    std::printf("status %d iters %d Time: %.3f ms\n", status, iters,
                std::chrono::duration<double, std::milli>(t1 - t0).count());
    const bool ok = status == MPC_STATUS_SOLVED && std::fabs(sol[1] + 12.782) < 0.01 &&
                    std::fabs(sol[2] - 42.606) < 0.01 && std::fabs(sol[8] - 42.606) < 0.01 &&
                    std::fabs(sol[5]) < 1e-3;
    // compute_grf + compute_joint_torques through the controller mirror: with the default
    // j_foot = I (A1CtrlStates.h:95) a stance leg gets tau = -f + gravity term, and the first
    // nine calls return zero torques (A1RobotControl.cpp:292-295)
    mpc_b200::A1RobotControl ctl;
    state.root_euler = {0, 0, 0};
    state.root_pos_d = state.root_pos;
    state.foot_pos_abs = foot_pos_rel;
    const std::array<double, 12> grf = ctl.compute_grf(state, dt);
    bool tok = true;
    for (int k = 0; k < 9; ++k) {
      ctl.compute_joint_torques(state);
      for (int i = 0; i < 12; ++i) tok = tok && state.joint_torques[i] == 0.0;
    }
    ctl.compute_joint_torques(state);
    for (int leg = 0; leg < 4; ++leg)
      for (int k = 0; k < 3; ++k) {
        const double want = state.contacts[leg] ? -grf[4 * k + leg] + state.torques_gravity[3 * leg + k]
                                                : state.torques_gravity[3 * leg + k];  // f_kin = 0
        tok = tok && std::fabs(state.joint_torques[3 * leg + k] - want) < 1e-4;
      }
    std::printf("torque map %s\n", tok ? "ok" : "MISMATCH");
    return (ok && tok) ? 0 : 1;
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "%s\n", ex.what());
    return 2;
  }
}
