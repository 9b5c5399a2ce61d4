// oracle/prep_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (the
// reference holds no expected outputs for these functions; see mpc_oracle.hpp).
//
// fp64 CPU restatement of what runs upstream of compute_grf every control tick:
//   quaternion -> rotation / euler / yaw rotation      GazeboA1ROS.cpp:262-269, utils/Utils.cpp:7-33
//   leg kinematics, foot positions and velocities      GazeboA1ROS.cpp:272-288 (A1Kinematics fk / jac;
//                                                       the generated code is restated from the
//                                                       hip-thigh-calf chain it encodes, rho_opt = 0)
//   root_ang_vel = R imu_ang_vel                        GazeboA1ROS.cpp:306
//   A1BasicEKF::init_state / update_estimation          A1BasicEKF.cpp:54-164
//   compute_walking_surface + terrain pitch             A1RobotControl.cpp:335-376, :566-582,
//                                                       utils/Utils.cpp:44-62, utils/filter.hpp:14-62
// The EKF's S^-1 products use Gaussian elimination with partial pivoting here (the reference calls
// fullPivHouseholderQr on the symmetric positive definite S; the device uses Cholesky).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <vector>

#include "../include/mpc_b200.h"

namespace prep_oracle {

struct RobotSlot {
  bool ekf_inited = false;
  double x[18];
  double P[18 * 18];
  std::deque<double> window;  // terrain_angle_filter, MovingWindowFilter(100)
  double sum = 0.0, correction = 0.0;
};

inline void neumaier(RobotSlot& s, double value) {  // filter.hpp:53-62
  const double new_sum = s.sum + value;
  if (std::abs(s.sum) >= std::abs(value)) s.correction += (s.sum - new_sum) + value;
  else s.correction += (value - new_sum) + s.sum;
  s.sum = new_sum;
}
inline double moving_average(RobotSlot& s, double value, size_t window_size = 100) {  // filter.hpp:26-39
  if (s.window.size() >= window_size) {
    neumaier(s, -s.window.front());
    s.window.pop_front();
  }
  neumaier(s, value);
  s.window.push_back(value);
  return (s.sum + s.correction) / double(window_size);
}

// p = o + Rx(q0) [ (0, d, 0) + Ry(q1) ( (0, 0, -lt) + Ry(q2) (0, 0, -lc) ) ],  J = dp/dq
inline void leg_fk_jac(const double* rho_fix, const double* q, double* p, double* J) {
  const double ox = rho_fix[0], oy = rho_fix[1], d = rho_fix[2], lt = rho_fix[3], lc = rho_fix[4];
  const double c0 = std::cos(q[0]), s0 = std::sin(q[0]), c1 = std::cos(q[1]), s1 = std::sin(q[1]);
  const double c12 = std::cos(q[1] + q[2]), s12 = std::sin(q[1] + q[2]);
  // point in the hip frame after the two pitch joints
  const double hx = -lt * s1 - lc * s12, hz = -lt * c1 - lc * c12;
  p[0] = ox + hx;
  p[1] = oy + d * c0 - hz * s0;
  p[2] = d * s0 + hz * c0;
  const double dhx1 = -lt * c1 - lc * c12, dhx2 = -lc * c12;  // d hx / d q1, d q2
  const double dhz1 = lt * s1 + lc * s12, dhz2 = lc * s12;    // d hz / d q1, d q2
  J[0] = 0.0;                 J[1] = dhx1;        J[2] = dhx2;
  J[3] = -d * s0 - hz * c0;   J[4] = -dhz1 * s0;  J[5] = -dhz2 * s0;
  J[6] = d * c0 - hz * s0;    J[7] = dhz1 * c0;   J[8] = dhz2 * c0;
}

// symmetric 3x3 pseudo-inverse through the eigen-decomposition (the SVD of a symmetric positive
// semi-definite matrix), tolerance eps * 3 * sigma_max (Utils.cpp:44-52)
inline void pinv_sym3(const double* M, double* out) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 9; ++i) A[i / 3][i % 3] = M[i];
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = i + 1; j < 3; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      if (A[p][q] == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) {
        const double akp = A[k][p], akq = A[k][q];
        A[k][p] = c * akp - s * akq;
        A[k][q] = s * akp + c * akq;
      }
      for (int k = 0; k < 3; ++k) {
        const double apk = A[p][k], aqk = A[q][k];
        A[p][k] = c * apk - s * aqk;
        A[q][k] = s * apk + c * aqk;
      }
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
      }
    }
  }
  double smax = 0;
  for (int i = 0; i < 3; ++i) smax = std::max(smax, std::abs(A[i][i]));
  const double tol = std::numeric_limits<double>::epsilon() * 3.0 * smax;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k)
        if (std::abs(A[k][k]) > tol) acc += V[i][k] * V[j][k] / A[k][k];
      out[3 * i + j] = acc;
    }
}

// solve S X = B in place (n x n, n x m), partial pivoting
inline void solve_general(std::vector<double> S, int n, std::vector<double>& B, int m) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int r = k + 1; r < n; ++r)
      if (std::abs(S[r * n + k]) > std::abs(S[piv * n + k])) piv = r;
    if (piv != k) {
      for (int c = 0; c < n; ++c) std::swap(S[k * n + c], S[piv * n + c]);
      for (int c = 0; c < m; ++c) std::swap(B[k * m + c], B[piv * m + c]);
    }
    for (int r = k + 1; r < n; ++r) {
      const double f = S[r * n + k] / S[k * n + k];
      for (int c = k; c < n; ++c) S[r * n + c] -= f * S[k * n + c];
      for (int c = 0; c < m; ++c) B[r * m + c] -= f * B[k * m + c];
    }
  }
  for (int r = n - 1; r >= 0; --r)
    for (int c = 0; c < m; ++c) {
      double acc = B[r * m + c];
      for (int k = r + 1; k < n; ++k) acc -= S[r * n + k] * B[k * m + c];
      B[r * m + c] = acc / S[r * n + r];
    }
}

inline void prep_tick(const PrepConfig& cfg, const RobotSensorIn& in, RobotSlot& slot, MpcStateIn& st, MpcTorqueIn& tin,
                      RobotPrepOut& ex) {
  std::memset(&st, 0, sizeof(st));
  std::memset(&tin, 0, sizeof(tin));
  std::memset(&ex, 0, sizeof(ex));
  // ---- orientation (Eigen::Quaterniond::toRotationMatrix, Utils::quat_to_euler) ----
  const double w = in.root_quat[0], x = in.root_quat[1], y = in.root_quat[2], z = in.root_quat[3];
  double R[9];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
  double euler[3];
  euler[0] = std::atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
  double sp = 2 * (w * y - z * x);
  sp = std::min(1.0, std::max(-1.0, sp));
  euler[1] = std::asin(sp);
  euler[2] = std::atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z));
  // ---- kinematics ----
  double prel[12], vrel[12], pabs[12], vabs[12];
  for (int leg = 0; leg < 4; ++leg) {
    double q[3], qd[3], p[3], J[9];
    for (int k = 0; k < 3; ++k) { q[k] = in.joint_pos[3 * leg + k]; qd[k] = in.joint_vel[3 * leg + k]; }
    leg_fk_jac(cfg.rho_fix + 5 * leg, q, p, J);
    for (int r = 0; r < 3; ++r) {
      prel[3 * leg + r] = p[r];
      vrel[3 * leg + r] = J[3 * r] * qd[0] + J[3 * r + 1] * qd[1] + J[3 * r + 2] * qd[2];
    }
    for (int r = 0; r < 3; ++r) {
      pabs[3 * leg + r] = R[3 * r] * prel[3 * leg] + R[3 * r + 1] * prel[3 * leg + 1] + R[3 * r + 2] * prel[3 * leg + 2];
      vabs[3 * leg + r] = R[3 * r] * vrel[3 * leg] + R[3 * r + 1] * vrel[3 * leg + 1] + R[3 * r + 2] * vrel[3 * leg + 2];
    }
    for (int k = 0; k < 9; ++k) tin.j_foot[9 * leg + k] = float(J[k]);
  }
  double wv[3];
  for (int r = 0; r < 3; ++r)
    wv[r] = R[3 * r] * in.imu_ang_vel[0] + R[3 * r + 1] * in.imu_ang_vel[1] + R[3 * r + 2] * in.imu_ang_vel[2];
  double root_pos[3] = {in.root_pos[0], in.root_pos[1], in.root_pos[2]};
  double root_vel[3] = {in.root_lin_vel[0], in.root_lin_vel[1], in.root_lin_vel[2]};
  // ---- A1BasicEKF ----
  double est_contacts[4] = {1, 1, 1, 1};
  if (cfg.use_estimator) {
    if (!slot.ekf_inited) {
      // init_state (:54-68): no update on this tick, root_pos stays the odometry value
      slot.ekf_inited = true;
      std::fill(slot.P, slot.P + 324, 0.0);
      for (int i = 0; i < 18; ++i) slot.P[19 * i] = 3.0;
      std::fill(slot.x, slot.x + 18, 0.0);
      slot.x[2] = 0.09;
      for (int leg = 0; leg < 4; ++leg)
        for (int r = 0; r < 3; ++r) slot.x[6 + 3 * leg + r] = pabs[3 * leg + r] + slot.x[r];
    } else {
      const double dt = in.dt;
      const int N = 18, M = 28;
      double u[3];
      for (int r = 0; r < 3; ++r)
        u[r] = R[3 * r] * in.imu_acc[0] + R[3 * r + 1] * in.imu_acc[1] + R[3 * r + 2] * in.imu_acc[2];
      u[2] += -9.81;
      for (int i = 0; i < 4; ++i)
        est_contacts[i] = (in.movement_mode == 0.0f) ? 1.0 : std::min(std::max(double(in.foot_force[i]) / 100.0, 0.0), 1.0);
      std::vector<double> A(N * N, 0.0), Q(N * N, 0.0), Rm(M * M, 0.0), C(M * N, 0.0);
      for (int i = 0; i < N; ++i) A[i * N + i] = 1.0;
      for (int i = 0; i < 3; ++i) A[i * N + 3 + i] = dt;
      for (int i = 0; i < 3; ++i) {
        Q[i * N + i] = 0.01 * dt / 20.0;
        Q[(3 + i) * N + 3 + i] = 0.01 * dt * 9.8 / 20.0;
      }
      for (int leg = 0; leg < 4; ++leg) {
        const double k = 1 + (1 - est_contacts[leg]) * 1e3;
        for (int r = 0; r < 3; ++r) {
          Q[(6 + 3 * leg + r) * N + 6 + 3 * leg + r] = k * dt * 0.01;
          Rm[(3 * leg + r) * M + 3 * leg + r] = k * 0.001;
          Rm[(12 + 3 * leg + r) * M + 12 + 3 * leg + r] = k * 0.1;
          C[(3 * leg + r) * N + r] = -1.0;
          C[(3 * leg + r) * N + 6 + 3 * leg + r] = 1.0;
          C[(12 + 3 * leg + r) * N + 3 + r] = 1.0;
        }
        Rm[(24 + leg) * M + 24 + leg] = cfg.assume_flat_ground ? k * 0.001 : 1e5;
        C[(24 + leg) * N + 6 + 3 * leg + 2] = 1.0;
      }
      // process update
      double xbar[18];
      for (int i = 0; i < N; ++i) xbar[i] = slot.x[i];
      for (int i = 0; i < 3; ++i) { xbar[i] += dt * slot.x[3 + i]; xbar[3 + i] += dt * u[i]; }
      std::vector<double> AP(N * N, 0.0), Pbar(N * N, 0.0);
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0;
          for (int k = 0; k < N; ++k) acc += A[i * N + k] * slot.P[k * N + j];
          AP[i * N + j] = acc;
        }
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0;
          for (int k = 0; k < N; ++k) acc += AP[i * N + k] * A[j * N + k];
          Pbar[i * N + j] = acc + Q[i * N + j];
        }
      // measurement
      double yv[28], yhat[28];
      for (int i = 0; i < M; ++i) {
        double acc = 0;
        for (int k = 0; k < N; ++k) acc += C[i * N + k] * xbar[k];
        yhat[i] = acc;
      }
      for (int leg = 0; leg < 4; ++leg) {
        const double* fk = prel + 3 * leg;
        const double wx = in.imu_ang_vel[0], wy = in.imu_ang_vel[1], wz = in.imu_ang_vel[2];
        const double cr[3] = {wy * fk[2] - wz * fk[1], wz * fk[0] - wx * fk[2], wx * fk[1] - wy * fk[0]};  // skew(w) fk
        double lv[3];
        for (int r = 0; r < 3; ++r) lv[r] = -vrel[3 * leg + r] - cr[r];
        for (int r = 0; r < 3; ++r) {
          yv[3 * leg + r] = pabs[3 * leg + r];
          const double rl = R[3 * r] * lv[0] + R[3 * r + 1] * lv[1] + R[3 * r + 2] * lv[2];
          yv[12 + 3 * leg + r] = (1.0 - est_contacts[leg]) * slot.x[3 + r] + est_contacts[leg] * rl;
        }
        yv[24 + leg] = (1.0 - est_contacts[leg]) * (slot.x[2] + fk[2]) + est_contacts[leg] * 0.0;
      }
      std::vector<double> CP(M * N, 0.0), S(M * M, 0.0);
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0;
          for (int k = 0; k < N; ++k) acc += C[i * N + k] * Pbar[k * N + j];
          CP[i * N + j] = acc;
        }
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
          double acc = 0;
          for (int k = 0; k < N; ++k) acc += CP[i * N + k] * C[j * N + k];
          S[i * M + j] = acc + Rm[i * M + j];
        }
      for (int i = 0; i < M; ++i)
        for (int j = i + 1; j < M; ++j) {
          const double v = 0.5 * (S[i * M + j] + S[j * M + i]);
          S[i * M + j] = S[j * M + i] = v;
        }
      std::vector<double> rhs(M * (1 + N));  // [error_y | C]
      for (int i = 0; i < M; ++i) {
        rhs[i * (1 + N)] = yv[i] - yhat[i];
        for (int j = 0; j < N; ++j) rhs[i * (1 + N) + 1 + j] = C[i * N + j];
      }
      solve_general(S, M, rhs, 1 + N);
      // x = xbar + Pbar C' S^-1 e ;  P = Pbar - Pbar C' S^-1 C Pbar
      for (int i = 0; i < N; ++i) {
        double acc = 0;
        for (int k = 0; k < M; ++k) acc += CP[k * N + i] * rhs[k * (1 + N)];  // (C Pbar)' = Pbar C'
        slot.x[i] = xbar[i] + acc;
      }
      std::vector<double> T(N * N, 0.0);  // Pbar C' (S^-1 C)
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0;
          for (int k = 0; k < M; ++k) acc += CP[k * N + i] * rhs[k * (1 + N) + 1 + j];
          T[i * N + j] = acc;
        }
      std::vector<double> Pn(N * N, 0.0);
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0;
          for (int k = 0; k < N; ++k) acc += T[i * N + k] * Pbar[k * N + j];
          Pn[i * N + j] = Pbar[i * N + j] - acc;
        }
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) slot.P[i * N + j] = 0.5 * (Pn[i * N + j] + Pn[j * N + i]);
      // reduce position drift (:143-148)
      if (slot.P[0] * slot.P[N + 1] - slot.P[1] * slot.P[N] > 1e-6) {
        for (int i = 0; i < 2; ++i)
          for (int j = 2; j < N; ++j) slot.P[i * N + j] = slot.P[j * N + i] = 0.0;
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) slot.P[i * N + j] /= 10.0;
      }
      for (int r = 0; r < 3; ++r) { root_pos[r] = slot.x[r]; root_vel[r] = slot.x[3 + r]; }
      for (int r = 0; r < 3; ++r) { ex.estimated_root_pos[r] = float(slot.x[r]); ex.estimated_root_vel[r] = float(slot.x[3 + r]); }
    }
  }
  // ---- terrain adaptation (compute_grf, MPC branch) ----
  double pitch_d = in.root_euler_d[1];
  double terrain_angle = 0.0;
  {
    // compute_walking_surface: a = pinv(W'W) W' z, W = [1, x, y]
    double WtW[9] = {0}, Wtz[3] = {0};
    for (int leg = 0; leg < 4; ++leg) {
      const double row[3] = {1.0, in.foot_pos_recent_contact[3 * leg], in.foot_pos_recent_contact[3 * leg + 1]};
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) WtW[3 * i + j] += row[i] * row[j];
        Wtz[i] += row[i] * in.foot_pos_recent_contact[3 * leg + 2];
      }
    }
    double Pi[9], a[3];
    pinv_sym3(WtW, Pi);
    for (int i = 0; i < 3; ++i) a[i] = Pi[3 * i] * Wtz[0] + Pi[3 * i + 1] * Wtz[1] + Pi[3 * i + 2] * Wtz[2];
    const double sc[3] = {a[1], a[2], -1.0};
    if (root_pos[2] > 0.1) {
      // cal_dihedral_angle against (0, 0, 1)
      const double cosang = std::abs(sc[2]) / std::sqrt(sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2]);
      terrain_angle = moving_average(slot, std::acos(cosang));
    }
    terrain_angle = std::min(0.5, std::max(-0.5, terrain_angle));
    const double frd = in.foot_pos_recent_contact[2] + in.foot_pos_recent_contact[5] - in.foot_pos_recent_contact[8] -
                       in.foot_pos_recent_contact[11];
    if (cfg.use_terrain_adapt) pitch_d = (frd > 0.05) ? -terrain_angle : terrain_angle;
  }
  // ---- pack ----
  for (int r = 0; r < 3; ++r) {
    st.euler[r] = float(euler[r]); st.pos[r] = float(root_pos[r]); st.ang_vel[r] = float(wv[r]);
    st.lin_vel[r] = float(root_vel[r]); st.euler_d[r] = in.root_euler_d[r];
    st.lin_vel_d[r] = in.root_lin_vel_d[r]; st.ang_vel_d[r] = in.root_ang_vel_d[r];
  }
  st.euler_d[1] = float(pitch_d);
  st.pos_d_z = in.root_pos_d_z;
  for (int i = 0; i < 9; ++i) st.rot_mat[i] = float(R[i]);
  for (int i = 0; i < 12; ++i) st.foot_pos_abs[i] = float(pabs[i]);
  for (int i = 0; i < 4; ++i) st.contacts[i] = in.contacts[i];
  for (int i = 0; i < 12; ++i) { tin.foot_forces_kin[i] = in.foot_forces_kin[i]; tin.torques_gravity[i] = float(cfg.torques_gravity[i]); }
  for (int i = 0; i < 3; ++i) tin.km_foot[i] = float(cfg.km_foot[i]);
  for (int r = 0; r < 3; ++r) { ex.root_euler[r] = float(euler[r]); ex.root_ang_vel[r] = float(wv[r]); }
  for (int i = 0; i < 9; ++i) ex.root_rot_mat[i] = float(R[i]);
  for (int i = 0; i < 12; ++i) {
    ex.foot_pos_rel[i] = float(prel[i]); ex.foot_vel_rel[i] = float(vrel[i]);
    ex.foot_pos_abs[i] = float(pabs[i]); ex.foot_vel_abs[i] = float(vabs[i]);
    ex.foot_pos_world[i] = float(pabs[i] + root_pos[i % 3]); ex.foot_vel_world[i] = float(vabs[i] + root_vel[i % 3]);
  }
  for (int i = 0; i < 4; ++i) ex.estimated_contacts[i] = float(est_contacts[i]);
  ex.terrain_pitch_angle = float(terrain_angle);
  ex.root_euler_d_pitch = float(pitch_d);
}

}  // namespace prep_oracle
