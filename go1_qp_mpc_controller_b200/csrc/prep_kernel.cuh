// Upstream state preparation on the device: one WARP per robot turns a
// RobotSensorIn record into the solver's MpcStateIn, the torque map's MpcTorqueIn and the
// derived quantities of RobotPrepOut, with the robot's estimator and terrain filter kept in HBM
// between ticks.  Restates
//   GazeboA1ROS.cpp:262-288, :306           orientation, leg kinematics, foot positions/velocities
//   legKinematics/A1Kinematics.cpp fk/jac   (the hip-thigh-calf chain, rho_opt = 0)
//   A1BasicEKF.cpp:54-164                   init_state / update_estimation (18 states, 28 measurements)
//   A1RobotControl.cpp:335-376, :566-582    terrain plane fit, moving-window pitch, root_euler_d[1]
// All arithmetic in f64; records are f32 like the rest of the boundary.  This is 0.6 % of the
// solver's flops: the design goal is "the batch never leaves the device", not peak rate.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"

namespace mpcb200 {

constexpr int kEkfN = 18, kEkfM = 28;
constexpr int kTerrainWindow = 100;  // MovingWindowFilter(100), A1RobotControl.cpp:52
// persistent slot (doubles): x | P | live | window ring | head | count | sum | correction
constexpr int kSlotX = 0, kSlotP = kEkfN, kSlotLive = kSlotP + kEkfN * kEkfN, kSlotWin = kSlotLive + 1,
              kSlotHead = kSlotWin + kTerrainWindow, kSlotCount = kSlotHead + 1, kSlotSum = kSlotCount + 1,
              kSlotCorr = kSlotSum + 1, kPrepSlotStride = 448;
static_assert(kSlotCorr < kPrepSlotStride, "slot layout");
constexpr int kPrepWarps = 2;  // robots per CTA

struct PrepParams {
  double rho_fix[20];
  double km_foot[3];
  double torques_gravity[12];
  int use_estimator, assume_flat_ground, use_terrain_adapt;
};

struct PrepWarpSmem {
  double Pbar[kEkfN * kEkfN];
  double CP[kEkfM * kEkfN];          // C Pbar
  double S[kEkfM * kEkfM];           // innovation covariance, then its Cholesky factor (lower)
  double RHS[kEkfM * (kEkfN + 1)];   // [y - yhat | C], then S^-1 [..]
  double T[kEkfN * kEkfN];           // Pbar C' S^-1 C
  double x[kEkfN], xbar[kEkfN];
  double prel[12], vrel[12], pabs[12], vabs[12];
  double R[9], misc[8];
};

// measurement matrix of A1BasicEKF (A1BasicEKF.cpp:10-17), element (i, j)
__device__ __forceinline__ double ekf_C(int i, int j) {
  if (i < 12) {
    const int leg = i / 3, r = i - 3 * leg;
    return (j == r) ? -1.0 : (j == 6 + 3 * leg + r) ? 1.0 : 0.0;
  }
  if (i < 24) {
    const int r = (i - 12) % 3;
    return (j == 3 + r) ? 1.0 : 0.0;
  }
  return (j == 6 + 3 * (i - 24) + 2) ? 1.0 : 0.0;
}

// (C X)_i for a length-18 column X given by an accessor: every row of C has one or two non-zeros
template <class F>
__device__ __forceinline__ double ekf_C_row(int i, F X) {
  if (i < 12) {
    const int leg = i / 3, r = i - 3 * leg;
    return X(6 + 3 * leg + r) - X(r);
  }
  if (i < 24) return X(3 + (i - 12) % 3);
  return X(6 + 3 * (i - 24) + 2);
}

__device__ __forceinline__ void leg_chain(const double* rho, const double* q, double* p, double* J) {
  const double ox = rho[0], oy = rho[1], d = rho[2], lt = rho[3], lc = rho[4];
  double s0, c0, s1, c1, s12, c12;
  sincos(q[0], &s0, &c0);
  sincos(q[1], &s1, &c1);
  sincos(q[1] + q[2], &s12, &c12);
  const double L = lt * c1 + lc * c12;   // extension below the hip axis
  const double X = -lt * s1 - lc * s12;  // forward reach
  p[0] = ox + X;
  p[1] = oy + d * c0 + L * s0;
  p[2] = d * s0 - L * c0;
  J[0] = 0.0;              J[1] = -L;      J[2] = -lc * c12;
  J[3] = -d * s0 + L * c0; J[4] = s0 * X;  J[5] = -s0 * lc * s12;
  J[6] = d * c0 + L * s0;  J[7] = -c0 * X; J[8] = c0 * lc * s12;
}

// pseudo-inverse of a symmetric 3x3 (Utils::pseudo_inverse on W'W): Jacobi eigen-decomposition,
// eigenvalues below eps * 3 * max are dropped
__device__ inline void pinv_sym3_dev(const double* M, double* out) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 9; ++i) A[i / 3][i % 3] = M[i];
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off < 1e-300) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      if (A[p][q] == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { const double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
      for (int k = 0; k < 3; ++k) { const double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
    }
  }
  const double smax = fmax(fabs(A[0][0]), fmax(fabs(A[1][1]), fabs(A[2][2])));
  const double tol = 2.220446049250313e-16 * 3.0 * smax;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k)
        if (fabs(A[k][k]) > tol) acc += V[i][k] * V[j][k] / A[k][k];
      out[3 * i + j] = acc;
    }
}

__global__ void __launch_bounds__(32 * kPrepWarps)
state_prep_kernel(const RobotSensorIn* __restrict__ sensors, int n, double* __restrict__ slots,
                  MpcStateIn* __restrict__ states, MpcTorqueIn* __restrict__ tin, RobotPrepOut* __restrict__ extras,
                  const __grid_constant__ PrepParams pp) {
  extern __shared__ __align__(16) unsigned char prep_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PrepWarpSmem& sm = reinterpret_cast<PrepWarpSmem*>(prep_smem_raw)[warp];
  const int p = blockIdx.x * kPrepWarps + warp;
  if (p >= n) return;  // warp-uniform
  const RobotSensorIn& in = sensors[p];
  double* slot = slots + size_t(p) * kPrepSlotStride;
  constexpr int N = kEkfN, M = kEkfM, W = kEkfN + 1;

  // ---- orientation: every lane redundantly (Quaterniond::toRotationMatrix, Utils::quat_to_euler) ----
  const double qw = in.root_quat[0], qx = in.root_quat[1], qy = in.root_quat[2], qz = in.root_quat[3];
  double R[9];
  R[0] = 1 - 2 * (qy * qy + qz * qz); R[1] = 2 * (qx * qy - qw * qz);     R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz);     R[4] = 1 - 2 * (qx * qx + qz * qz); R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy);     R[7] = 2 * (qy * qz + qw * qx);     R[8] = 1 - 2 * (qx * qx + qy * qy);
  double euler[3];
  euler[0] = atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
  euler[1] = asin(fmin(1.0, fmax(-1.0, 2 * (qw * qy - qz * qx))));
  euler[2] = atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
  const double iw[3] = {(double)in.imu_ang_vel[0], (double)in.imu_ang_vel[1], (double)in.imu_ang_vel[2]};
  double wv[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) wv[r] = R[3 * r] * iw[0] + R[3 * r + 1] * iw[1] + R[3 * r + 2] * iw[2];

  // ---- kinematics: lane = leg ----
  if (lane < 4) {
    const int leg = lane;
    double q[3], qd[3], pl[3], J[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { q[k] = (double)in.joint_pos[3 * leg + k]; qd[k] = (double)in.joint_vel[3 * leg + k]; }
    leg_chain(pp.rho_fix + 5 * leg, q, pl, J);
    double vl[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) vl[r] = J[3 * r] * qd[0] + J[3 * r + 1] * qd[1] + J[3 * r + 2] * qd[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sm.prel[3 * leg + r] = pl[r];
      sm.vrel[3 * leg + r] = vl[r];
      sm.pabs[3 * leg + r] = R[3 * r] * pl[0] + R[3 * r + 1] * pl[1] + R[3 * r + 2] * pl[2];
      sm.vabs[3 * leg + r] = R[3 * r] * vl[0] + R[3 * r + 1] * vl[1] + R[3 * r + 2] * vl[2];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) tin[p].j_foot[9 * leg + k] = (float)J[k];
  }
  __syncwarp();

  double root_pos[3] = {(double)in.root_pos[0], (double)in.root_pos[1], (double)in.root_pos[2]};
  double root_vel[3] = {(double)in.root_lin_vel[0], (double)in.root_lin_vel[1], (double)in.root_lin_vel[2]};
  double estc[4] = {1.0, 1.0, 1.0, 1.0};
  bool have_estimate = false;

  // ---- A1BasicEKF ----
  if (pp.use_estimator) {
    const bool live = slot[kSlotLive] != 0.0;
    __syncwarp();
    if (!live) {
      // init_state: x = (0, 0, 0.09, 0, 0, 0, R fk + pos), P = 3 I; no update on this tick
      for (int i = lane; i < N * N; i += 32) slot[kSlotP + i] = (i / N == i % N) ? 3.0 : 0.0;
      if (lane < N) {
        double v = 0.0;
        if (lane == 2) v = 0.09;
        if (lane >= 6) v = sm.pabs[lane - 6] + (((lane - 6) % 3 == 2) ? 0.09 : 0.0);
        slot[kSlotX + lane] = v;
      }
      if (lane == 0) slot[kSlotLive] = 1.0;
    } else {
      const double dt = (double)in.dt;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        estc[i] = (in.movement_mode == 0.0f) ? 1.0 : fmin(fmax((double)in.foot_force[i] / 100.0, 0.0), 1.0);
      double u[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        u[r] = R[3 * r] * (double)in.imu_acc[0] + R[3 * r + 1] * (double)in.imu_acc[1] + R[3 * r + 2] * (double)in.imu_acc[2];
      u[2] += -9.81;
      if (lane < N) sm.x[lane] = slot[kSlotX + lane];
      __syncwarp();
      if (lane < N) {
        double v = sm.x[lane];
        if (lane < 3) v += dt * sm.x[3 + lane];
        else if (lane < 6) v += dt * u[lane - 3];
        sm.xbar[lane] = v;
      }
      // Pbar = A P A' + Q with A = I + dt E(0:3, 3:6): rows/cols 0..2 pick up dt times rows/cols 3..5
      for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - N * i;
        const double* P = slot + kSlotP;
        double v = P[i * N + j];
        if (i < 3) v += dt * P[(3 + i) * N + j];
        if (j < 3) v += dt * P[i * N + 3 + j];
        if (i < 3 && j < 3) v += dt * dt * P[(3 + i) * N + 3 + j];
        if (i == j) {
          double qd;
          if (i < 3) qd = 0.01 * dt / 20.0;
          else if (i < 6) qd = 0.01 * dt * 9.8 / 20.0;
          else qd = (1.0 + (1.0 - estc[(i - 6) / 3]) * 1e3) * dt * 0.01;
          v += qd;
        }
        sm.Pbar[idx] = v;
      }
      __syncwarp();
      // CP = C Pbar
      for (int idx = lane; idx < M * N; idx += 32) {
        const int i = idx / N, j = idx - N * i;
        sm.CP[idx] = ekf_C_row(i, [&](int k) { return sm.Pbar[k * N + j]; });
      }
      __syncwarp();
      // S = sym(CP C' + R)
      for (int idx = lane; idx < M * M; idx += 32) {
        const int i = idx / M, j = idx - M * i;
        const double acc = ekf_C_row(j, [&](int k) { return sm.CP[i * N + k]; });
        const double accT = ekf_C_row(i, [&](int k) { return sm.CP[j * N + k]; });
        double v = 0.5 * (acc + accT);
        if (i == j) {
          double rd;
          if (i < 12) rd = (1.0 + (1.0 - estc[i / 3]) * 1e3) * 0.001;
          else if (i < 24) rd = (1.0 + (1.0 - estc[(i - 12) / 3]) * 1e3) * 0.1;
          else rd = pp.assume_flat_ground ? (1.0 + (1.0 - estc[i - 24]) * 1e3) * 0.001 : 1e5;
          v += rd;
        }
        sm.S[idx] = v;
      }
      // right-hand sides [y - C xbar | C]
      for (int idx = lane; idx < M * W; idx += 32) {
        const int i = idx / W, j = idx - W * i;
        double v;
        if (j > 0) {
          v = ekf_C(i, j - 1);
        } else {
          double yv;
          if (i < 12) {
            yv = sm.pabs[i];
          } else if (i < 24) {
            const int leg = (i - 12) / 3, r = (i - 12) - 3 * leg;
            const double* fk = sm.prel + 3 * leg;
            const double cr[3] = {iw[1] * fk[2] - iw[2] * fk[1], iw[2] * fk[0] - iw[0] * fk[2], iw[0] * fk[1] - iw[1] * fk[0]};
            const double l0 = -sm.vrel[3 * leg] - cr[0], l1 = -sm.vrel[3 * leg + 1] - cr[1], l2 = -sm.vrel[3 * leg + 2] - cr[2];
            const double rl = R[3 * r] * l0 + R[3 * r + 1] * l1 + R[3 * r + 2] * l2;
            yv = (1.0 - estc[leg]) * sm.x[3 + r] + estc[leg] * rl;
          } else {
            const int leg = i - 24;
            yv = (1.0 - estc[leg]) * (sm.x[2] + sm.prel[3 * leg + 2]);
          }
          v = yv - ekf_C_row(i, [&](int k) { return sm.xbar[k]; });
        }
        sm.RHS[idx] = v;
      }
      __syncwarp();
      // Cholesky S = L L' (left-looking, lanes over rows), then S^-1 RHS by two triangular solves
      for (int k = 0; k < M; ++k) {
        if (lane == 0) {
          double d = sm.S[k * M + k];
          for (int j = 0; j < k; ++j) d -= sm.S[k * M + j] * sm.S[k * M + j];
          sm.S[k * M + k] = sqrt(d);
        }
        __syncwarp();
        const double dk = sm.S[k * M + k];
        for (int r = k + 1 + lane; r < M; r += 32) {
          double v = sm.S[r * M + k];
          for (int j = 0; j < k; ++j) v -= sm.S[r * M + j] * sm.S[k * M + j];
          sm.S[r * M + k] = v / dk;
        }
        __syncwarp();
      }
      if (lane < W) {
        const int c = lane;
        for (int i = 0; i < M; ++i) {
          double v = sm.RHS[i * W + c];
          for (int j = 0; j < i; ++j) v -= sm.S[i * M + j] * sm.RHS[j * W + c];
          sm.RHS[i * W + c] = v / sm.S[i * M + i];
        }
        for (int i = M - 1; i >= 0; --i) {
          double v = sm.RHS[i * W + c];
          for (int j = i + 1; j < M; ++j) v -= sm.S[j * M + i] * sm.RHS[j * W + c];
          sm.RHS[i * W + c] = v / sm.S[i * M + i];
        }
      }
      __syncwarp();
      // x = xbar + (C Pbar)' S^-1 e ;  T = (C Pbar)' S^-1 C
      if (lane < N) {
        double acc = 0.0;
        for (int k = 0; k < M; ++k) acc += sm.CP[k * N + lane] * sm.RHS[k * W];
        const double v = sm.xbar[lane] + acc;
        slot[kSlotX + lane] = v;
        sm.x[lane] = v;
      }
      for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - N * i;
        double acc = 0.0;
        for (int k = 0; k < M; ++k) acc += sm.CP[k * N + i] * sm.RHS[k * W + 1 + j];
        sm.T[idx] = acc;
      }
      __syncwarp();
      // P = sym(Pbar - T Pbar), staged in CP (free now)
      double* Pn = sm.CP;
      for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - N * i;
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc += sm.T[i * N + k] * sm.Pbar[k * N + j];
        Pn[idx] = sm.Pbar[idx] - acc;
      }
      __syncwarp();
      const double p00 = 0.5 * (Pn[0] + Pn[0]), p01 = 0.5 * (Pn[1] + Pn[N]), p11 = Pn[N + 1];
      const bool damp = (p00 * p11 - p01 * p01) > 1e-6;  // reduce position drift (:143-148)
      for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - N * i;
        double v = 0.5 * (Pn[i * N + j] + Pn[j * N + i]);
        if (damp) {
          if ((i < 2) != (j < 2)) v = 0.0;
          else if (i < 2 && j < 2) v /= 10.0;
        }
        slot[kSlotP + idx] = v;
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 3; ++r) { root_pos[r] = sm.x[r]; root_vel[r] = sm.x[3 + r]; }
      have_estimate = true;
    }
  }

  // ---- terrain adaptation and packing: lane 0 ----
  if (lane == 0) {
    const float* frc = in.foot_pos_recent_contact;
    double WtW[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Wtz[3] = {0, 0, 0};
    for (int leg = 0; leg < 4; ++leg) {
      const double row[3] = {1.0, (double)frc[3 * leg], (double)frc[3 * leg + 1]};
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) WtW[3 * i + j] += row[i] * row[j];
        Wtz[i] += row[i] * (double)frc[3 * leg + 2];
      }
    }
    double Pi[9];
    pinv_sym3_dev(WtW, Pi);
    const double a1 = Pi[3] * Wtz[0] + Pi[4] * Wtz[1] + Pi[5] * Wtz[2];
    const double a2 = Pi[6] * Wtz[0] + Pi[7] * Wtz[1] + Pi[8] * Wtz[2];
    double terrain_angle = 0.0;
    if (root_pos[2] > 0.1) {
      // dihedral angle between the fitted plane (a1, a2, -1) and flat ground (0, 0, 1), through
      // the moving-window filter (Neumaier sum, filter.hpp:26-62)
      const double value = acos(1.0 / sqrt(a1 * a1 + a2 * a2 + 1.0));
      int head = (int)slot[kSlotHead], count = (int)slot[kSlotCount];
      double sum = slot[kSlotSum], corr = slot[kSlotCorr];
      auto neumaier = [&](double v) {
        const double ns = sum + v;
        if (fabs(sum) >= fabs(v)) corr += (sum - ns) + v;
        else corr += (v - ns) + sum;
        sum = ns;
      };
      if (count >= kTerrainWindow) neumaier(-slot[kSlotWin + head]);  // the oldest value sits at head
      else ++count;
      neumaier(value);
      slot[kSlotWin + head] = value;
      head = (head + 1 == kTerrainWindow) ? 0 : head + 1;
      slot[kSlotHead] = (double)head;
      slot[kSlotCount] = (double)count;
      slot[kSlotSum] = sum;
      slot[kSlotCorr] = corr;
      terrain_angle = (sum + corr) / (double)kTerrainWindow;
    }
    terrain_angle = fmin(0.5, fmax(-0.5, terrain_angle));
    const double frd = (double)frc[2] + (double)frc[5] - (double)frc[8] - (double)frc[11];
    double pitch_d = (double)in.root_euler_d[1];
    if (pp.use_terrain_adapt) pitch_d = (frd > 0.05) ? -terrain_angle : terrain_angle;

    MpcStateIn& st = states[p];
    for (int r = 0; r < 3; ++r) {
      st.euler[r] = (float)euler[r];
      st.pos[r] = (float)root_pos[r];
      st.ang_vel[r] = (float)wv[r];
      st.lin_vel[r] = (float)root_vel[r];
      st.euler_d[r] = in.root_euler_d[r];
      st.lin_vel_d[r] = in.root_lin_vel_d[r];
      st.ang_vel_d[r] = in.root_ang_vel_d[r];
    }
    st.euler_d[1] = (float)pitch_d;
    st.pos_d_z = in.root_pos_d_z;
    for (int i = 0; i < 9; ++i) st.rot_mat[i] = (float)R[i];
    for (int i = 0; i < 12; ++i) st.foot_pos_abs[i] = (float)sm.pabs[i];
    for (int i = 0; i < 4; ++i) st.contacts[i] = in.contacts[i];
    st.pad = 0.0f;
    MpcTorqueIn& t = tin[p];
    for (int i = 0; i < 12; ++i) { t.foot_forces_kin[i] = in.foot_forces_kin[i]; t.torques_gravity[i] = (float)pp.torques_gravity[i]; }
    for (int i = 0; i < 3; ++i) t.km_foot[i] = (float)pp.km_foot[i];
    t.pad = 0.0f;
    if (extras != nullptr) {
      RobotPrepOut& ex = extras[p];
      for (int r = 0; r < 3; ++r) {
        ex.root_euler[r] = (float)euler[r];
        ex.root_ang_vel[r] = (float)wv[r];
        ex.estimated_root_pos[r] = have_estimate ? (float)root_pos[r] : 0.0f;
        ex.estimated_root_vel[r] = have_estimate ? (float)root_vel[r] : 0.0f;
      }
      for (int i = 0; i < 9; ++i) ex.root_rot_mat[i] = (float)R[i];
      for (int i = 0; i < 12; ++i) {
        ex.foot_pos_rel[i] = (float)sm.prel[i];
        ex.foot_vel_rel[i] = (float)sm.vrel[i];
        ex.foot_pos_abs[i] = (float)sm.pabs[i];
        ex.foot_vel_abs[i] = (float)sm.vabs[i];
        ex.foot_pos_world[i] = (float)(sm.pabs[i] + root_pos[i % 3]);
        ex.foot_vel_world[i] = (float)(sm.vabs[i] + root_vel[i % 3]);
      }
      for (int i = 0; i < 4; ++i) ex.estimated_contacts[i] = (float)estc[i];
      ex.terrain_pitch_angle = (float)terrain_angle;
      ex.root_euler_d_pitch = (float)pitch_d;
      for (int i = 0; i < 29; ++i) ex.pad[i] = 0.0f;
    }
  }
}

}  // namespace mpcb200
