// Drives mpc_b200::A1RobotControl::compute_grf through consecutive control ticks of ONE robot, the way
// thread 1 of the reference calls it (MainGazebo.cpp:47-81 -> GazeboA1ROS.cpp:112-115 ->
// A1RobotControl.cpp:321): the controller keeps one solver alive, so tick 0 is initSolver and every
// later tick is updateHessianMatrix / updateGradient / updateBounds + a warm solve (:522-540).
// Input: a file of MpcStateIn records (one per tick).  Output: one line per tick,
//   tick status iters grf[12 leg-major]
// which tests/test_cpp_shim.py compares with the oracle's MpcStream tick by tick.
#include <cstdio>
#include <vector>

#include "convex_mpc_b200.hpp"

using namespace mpc_b200;

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s records.bin [hardware]\n", argv[0]); return 2; }
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  std::vector<MpcStateIn> recs;
  MpcStateIn r;
  while (std::fread(&r, sizeof(r), 1, f) == 1) recs.push_back(r);
  std::fclose(f);
  MpcConfig cfg;
  if (argc > 2) mpc_config_hardware(&cfg); else mpc_config_default(&cfg);
  try {
    A1RobotControl ctl;
    A1CtrlStates s;
    s.robot_mass = cfg.mass;
    for (int i = 0; i < 9; ++i) s.a1_trunk_inertia[i] = cfg.inertia[i];
    for (int i = 0; i < 13; ++i) s.q_weights[i] = cfg.q_weights[i];
    for (int i = 0; i < 12; ++i) s.r_weights[i] = cfg.r_weights[i];
    for (size_t t = 0; t < recs.size(); ++t) {
      const MpcStateIn& in = recs[t];
      for (int i = 0; i < 3; ++i) {
        s.root_euler[i] = in.euler[i]; s.root_pos[i] = in.pos[i];
        s.root_ang_vel[i] = in.ang_vel[i]; s.root_lin_vel[i] = in.lin_vel[i];
        s.root_euler_d[i] = in.euler_d[i]; s.root_lin_vel_d[i] = in.lin_vel_d[i];
        s.root_ang_vel_d[i] = in.ang_vel_d[i];
      }
      s.root_pos_d = {0.0, 0.0, in.pos_d_z};
      for (int i = 0; i < 9; ++i) s.root_rot_mat[i] = in.rot_mat[i];
      for (int leg = 0; leg < 4; ++leg) {
        for (int k = 0; k < 3; ++k) s.foot_pos_abs[4 * k + leg] = in.foot_pos_abs[3 * leg + k];
        s.contacts[leg] = in.contacts[leg] != 0.0f;
      }
      const std::array<double, 12> g = ctl.compute_grf(s, 0.0025);
      std::printf("%zu %d %d", t, ctl.last_status, ctl.last_iters);
      for (int leg = 0; leg < 4; ++leg)
        for (int k = 0; k < 3; ++k) std::printf(" %.9g", g[4 * k + leg]);
      std::printf("\n");
    }
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "%s\n", ex.what());
    return 2;
  }
  return 0;
}
