// Torque map of compute_joint_torques (A1RobotControl.cpp:289-319), SURVEY.md 8f row 3.
//   stance leg : tau = J^T (-f_grf)                                  (:301-303)
//   swing leg  : J tau = km .* f_kin, 3x3 LU with partial pivoting    (:304-308, Eigen lu().solve)
//   + torques_gravity                                                 (:311)
//   NaN components are flagged; the caller keeps its previous value  (:314-317)
// leg_torque() is shared by the fused epilogue of admm_solve_kernel (H = 10, cold and warm) and
// by torque_map_kernel, which serves the long-horizon and stance-balance engines from the
// written results.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"

namespace mpcb200 {

// J row-major 3x3 (foot velocity = J * joint velocity); returns the three joint torques of one leg
__device__ __forceinline__ void leg_torque(const float* __restrict__ J, bool contact, double fx, double fy, double fz,
                                           const float* __restrict__ fkin, const float* __restrict__ km,
                                           const float* __restrict__ grav, double (&tau)[3]) {
  double a[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) a[r][c] = (double)J[3 * r + c];
  if (contact) {
#pragma unroll
    for (int k = 0; k < 3; ++k) tau[k] = -(a[0][k] * fx + a[1][k] * fy + a[2][k] * fz);
  } else {
    double b[3] = {(double)km[0] * (double)fkin[0], (double)km[1] * (double)fkin[1], (double)km[2] * (double)fkin[2]};
    // Gaussian elimination with row pivoting on the largest magnitude (PartialPivLU); a singular
    // Jacobian divides by zero and the NaN/inf go to the guard, as in the reference
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int piv = k;
      double best = fabs(a[k][k]);
#pragma unroll
      for (int r = k + 1; r < 3; ++r) {
        const double v = fabs(a[r][k]);
        if (v > best) { best = v; piv = r; }
      }
#pragma unroll
      for (int r = k + 1; r < 3; ++r) {
        if (piv == r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { const double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
          const double t = b[k]; b[k] = b[r]; b[r] = t;
        }
      }
#pragma unroll
      for (int r = k + 1; r < 3; ++r) {
        const double l = a[r][k] / a[k][k];
#pragma unroll
        for (int c = k + 1; c < 3; ++c) a[r][c] -= l * a[k][c];
        b[r] -= l * b[k];
      }
    }
    tau[2] = b[2] / a[2][2];
    tau[1] = (b[1] - a[1][2] * tau[2]) / a[1][1];
    tau[0] = (b[0] - a[0][1] * tau[1] - a[0][2] * tau[2]) / a[0][0];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) tau[k] += (double)grav[k];
}

// One thread per (robot, leg) over results already in memory.  contacts: float flags, `stride`
// floats between robots, first flag at `offset` (MpcStateIn: 48 / 43, BalanceStateIn: 64 / 54).
__global__ void torque_map_kernel(const MpcResult* __restrict__ results, const float* __restrict__ state_words,
                                  int stride, int offset, const MpcTorqueIn* __restrict__ tin,
                                  MpcTorqueOut* __restrict__ tout, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = idx >> 2, leg = idx & 3;
  unsigned nanbits = 0;
  if (p < n) {
    const MpcTorqueIn& t = tin[p];
    const float* f = results[p].grf + 3 * leg;
    const bool contact = state_words[size_t(p) * stride + offset + leg] != 0.0f;
    double tau[3];
    leg_torque(t.j_foot + 9 * leg, contact, (double)f[0], (double)f[1], (double)f[2], t.foot_forces_kin + 3 * leg,
               t.km_foot, t.torques_gravity + 3 * leg, tau);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool bad = isnan(tau[k]);
      tout[p].joint_torques[3 * leg + k] = bad ? 0.0f : (float)tau[k];
      if (bad) nanbits |= 1u << (3 * leg + k);
    }
  }
  // the four legs of a robot sit in one aligned quad of lanes
  nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 1);
  nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 2);
  if (p < n && leg == 0) tout[p].nan_mask = (int32_t)nanbits;
}

}  // namespace mpcb200
