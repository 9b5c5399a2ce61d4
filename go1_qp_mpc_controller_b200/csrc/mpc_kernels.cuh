// Device kernels of the batched convex-MPC GRF engine (sm_100a).
//
//   qp_build_kernel     K0+K1+K2  state record -> A_d, B_d, A_qp powers, block-lower-triangular
//                                 B_qp, Hessian B'QB+R, gradient, bounds
//                                 (ConvexMpc.cpp:110-245, A1RobotControl.cpp:452-518)
//   admm_solve_kernel   K3+K4+K5  OSQP-equivalent ADMM: Ruiz scaling, K = P+sigma I+A'rho A,
//                                 K^-1 by symmetric sweep held in REGISTERS, ADMM loop,
//                                 residual termination, rho adaptation, unscale + R'f writer
//                                 (OSQP 0.6.x as driven by A1RobotControl.cpp:522-561)
//
// Precision (measured on the CPU arithmetic model, see DESIGN.md "precision"):
// the GRF parity gate (1e-3 vs the fp64 oracle at eps 1e-5) needs the Hessian
// ACCUMULATED in fp64 and K, K^-1 and the ADMM iterates in fp64; fp32 is fine
// for the QP as read back through mpc_get_qp.  Between the two kernels P and q
// stay in f64 (HBM traffic is ~1 GB per 4096-state step, irrelevant next to the
// solve), which keeps the device iterate sequence identical to the oracle's:
// with f32 hand-over 0.3 % of states flipped a termination check.  B200 runs
// DFMA at half the FFMA rate (measured 17.1 T DFMA/s).
//
// Thread layout shared by both kernels (H = 10, n = 120): 480 threads =
// 30 row groups x 16 column groups; thread (rg, cg) owns the 4 x 8 register
// tile rows 4rg..4rg+3 x columns {32i + 2cg, 32i + 2cg + 1 : i = 0..3}.
// Columns are interleaved in pairs at stride 32 so that one LDS.128 per i
// fetches a thread's column pair and the 16 lanes of a half-warp read 256
// contiguous bytes (no bank conflicts); the other half-warp (next row group)
// reads the same addresses (broadcast).  Row sums are finished with a
// reduce-scatter over the 16 lanes (5 double shuffles for 4 rows).  Matrices
// are padded to 128 columns; columns 120..127 are structurally zero.
// v1 of this file used a 1 x 30 tile: 4-way bank conflicts and 3x the operand
// traffic (profiles/r01_v1_admm_solve_ncu_summary.txt).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"

namespace mpcb200 {

constexpr int kH = 10;
constexpr int kN = 12 * kH;  // 120 variables
constexpr int kS = 13 * kH;  // 130 stacked states
constexpr int kM = 20 * kH;  // 200 constraint rows
constexpr int kLegSteps = 4 * kH;
constexpr int kThreads = 480;
constexpr int kNP = 128;     // padded column count
constexpr int kWarps = kThreads / 32;

// float offsets inside MpcStateIn
constexpr int kOffEuler = 0, kOffPos = 3, kOffAngVel = 6, kOffLinVel = 9, kOffEulerD = 12,
              kOffPosDz = 15, kOffLinVelD = 16, kOffAngVelD = 19, kOffRot = 22, kOffFoot = 31,
              kOffContacts = 43;

struct BuildParams {
  double dt, mu, fz_min, fz_max, mass;
  double inertia[9];
  double Qd[13];  // 2 * q_weights (ConvexMpc.cpp:20)
  double Rd[12];  // 2 * r_weights (ConvexMpc.cpp:41)
  int exact_discretization, foot_drift, gait_aware;  // SURVEY.md 8f row 4, all 0 = the reference
};

struct SolveParams {
  double rho, sigma, alpha, eps_abs, eps_rel, adaptive_rho_tolerance, mu;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
};

// column owned by column group cg at tile position jj (0..7)
__host__ __device__ constexpr int tile_col(int cg, int jj) { return 32 * (jj >> 1) + 2 * cg + (jj & 1); }

// Reduce 4 per-row partials over the 16 column-group lanes of a half-warp.
// Returns the total of row `owned_row(cg)`; lanes cg, cg+1, cg+2, cg+3 (cg%4==0)
// all hold the same value.
__device__ __forceinline__ int owned_row(int cg) { return 2 * ((cg >> 3) & 1) + ((cg >> 2) & 1); }

__device__ __forceinline__ double reduce_scatter_sum(const double (&s)[4], int cg) {
  const bool h8 = (cg & 8) != 0, h4 = (cg & 4) != 0;
  double k0 = h8 ? s[2] : s[0], k1 = h8 ? s[3] : s[1];
  const double t0 = h8 ? s[0] : s[2], t1 = h8 ? s[1] : s[3];
  k0 += __shfl_xor_sync(0xffffffffu, t0, 8);
  k1 += __shfl_xor_sync(0xffffffffu, t1, 8);
  double k = h4 ? k1 : k0;
  const double t = h4 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, t, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
__device__ __forceinline__ double reduce_scatter_max(const double (&s)[4], int cg) {
  const bool h8 = (cg & 8) != 0, h4 = (cg & 4) != 0;
  double k0 = h8 ? s[2] : s[0], k1 = h8 ? s[3] : s[1];
  const double t0 = h8 ? s[0] : s[2], t1 = h8 ? s[1] : s[3];
  k0 = fmax(k0, __shfl_xor_sync(0xffffffffu, t0, 8));
  k1 = fmax(k1, __shfl_xor_sync(0xffffffffu, t1, 8));
  double k = h4 ? k1 : k0;
  const double t = h4 ? k0 : k1;
  k = fmax(k, __shfl_xor_sync(0xffffffffu, t, 4));
  k = fmax(k, __shfl_xor_sync(0xffffffffu, k, 2));
  k = fmax(k, __shfl_xor_sync(0xffffffffu, k, 1));
  return k;
}

// ---------------------------------------------------------------------------
// K0+K1+K2: QP build.  One CTA per problem, grid-stride over problems.
// ---------------------------------------------------------------------------
struct BuildSmem {
  double Bq[kS * kNP];           // B_qp, row stride 128 (133,120 B); cols 120..127 zero
  double Apow[(kH + 1) * 169];   // A_d^0 .. A_d^H
  double Bd[kH * 156];           // B_mat_d_list
  double xref[kS];
  double tmp[kS];
  double x0[16];
  float st[48];
  int contacts[kH * 4];  // per step and leg (replicated unless the horizon is gait aware)
};

// Optional caller-written model for the ConvexMpc surface (one problem each):
// A_mat_d 13x13, B_mat_d_list 13H x 12, mpc_states 13, mpc_states_d 13H, contacts 4.
struct ModelIn {
  const double* A_d;
  const double* B_d_list;
  const double* x0;
  const double* x_ref;
  const int* contacts;
};

__global__ void __launch_bounds__(kThreads, 1)
qp_build_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait, ModelIn model, int num,
                double* __restrict__ P_out,
                double* __restrict__ q_out, float* __restrict__ l_out, float* __restrict__ u_out,
                const __grid_constant__ BuildParams bp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BuildSmem& sm = *reinterpret_cast<BuildSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int rg = tid >> 4, cg = tid & 15;

  for (int p = blockIdx.x; p < num; p += gridDim.x) {
    __syncthreads();  // smem reuse across problems
    if (states != nullptr) {
      // ---- K0: coalesced record load (48 floats) ----
      if (tid < 48) sm.st[tid] = reinterpret_cast<const float*>(states + p)[tid];
      __syncthreads();
      const float* st = sm.st;
      // ---- A_d = I + dt*A_c (ConvexMpc.cpp:110-130, :149-150) ----
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        const double yaw = (double)st[kOffEuler + 2];
        double s, c;
        sincos(yaw, &s, &c);
        double a = 0.0;
        if (rr == 0 && cc == 6) a = c;
        if (rr == 0 && cc == 7) a = s;
        if (rr == 1 && cc == 6) a = -s;
        if (rr == 1 && cc == 7) a = c;
        if (rr == 2 && cc == 8) a = 1.0;
        if (rr >= 3 && rr <= 5 && cc == rr + 6) a = 1.0;
        if (rr == 11 && cc == 12) a = 1.0;
        // exact discretisation: A_c^3 = 0 and A_c^2 has the single entry (5, 12) = 1
        if (bp.exact_discretization && rr == 5 && cc == 12) a = 0.5 * bp.dt;
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = ((rr == cc) ? 1.0 : 0.0) + a * bp.dt;
      } else if (tid >= 192 && tid < 192 + 156) {
        for (int i = 0; i < kH; ++i) sm.Bd[i * 156 + tid - 192] = 0.0;
      } else if (tid >= 352 && tid < 352 + 13) {
        // mpc_states (A1RobotControl.cpp:452-456)
        const int k = tid - 352;
        sm.x0[k] = (k < 12) ? (double)st[k] : -9.8;
      } else if (tid >= 384 && tid < 384 + kH) {
        // mpc_states_d step i (A1RobotControl.cpp:470-488)
        const int i = tid - 384;
        const double R0 = st[kOffRot + 0], R1 = st[kOffRot + 1], R2 = st[kOffRot + 2];
        const double R3 = st[kOffRot + 3], R4 = st[kOffRot + 4], R5 = st[kOffRot + 5];
        const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
        const double vwx = R0 * vx + R1 * vy + R2 * vz;
        const double vwy = R3 * vx + R4 * vy + R5 * vz;
        const double dt = bp.dt;
        double* d = &sm.xref[13 * i];
        d[0] = (double)st[kOffEulerD];
        d[1] = (double)st[kOffEulerD + 1];
        d[2] = (double)st[kOffEuler + 2] + (double)st[kOffAngVelD + 2] * dt * (double)(i + 1);
        d[3] = (double)st[kOffPos] + vwx * dt * (double)(i + 1);
        d[4] = (double)st[kOffPos + 1] + vwy * dt * (double)(i + 1);
        d[5] = (double)st[kOffPosDz];
        d[6] = (double)st[kOffAngVelD];
        d[7] = (double)st[kOffAngVelD + 1];
        d[8] = (double)st[kOffAngVelD + 2];
        d[9] = vwx;
        d[10] = vwy;
        d[11] = 0.0;
        d[12] = -9.8;
      } else if (tid >= 416 && tid < 416 + 4 * kH) {
        const int i = (tid - 416) >> 2, leg = (tid - 416) & 3;
        int c = st[kOffContacts + leg] != 0.0f;
        if (bp.gait_aware && i > 0) {
          // planned contact of step i from the gait counter (A1RobotControl.cpp:156-164)
          const float* g = reinterpret_cast<const float*>(gait + p);
          const double cnt = fmod((double)g[leg] + (double)i * (double)g[10] * (double)g[4 + leg], (double)g[8]);
          c = cnt <= (double)g[9];
        }
        sm.contacts[4 * i + leg] = c;
      }
      __syncthreads();
      // ---- B_d = dt*B_c, one thread per (step, leg) (ConvexMpc.cpp:132-143, :151); without
      //      foot_drift only step 0 is computed and then copied (A1RobotControl.cpp:498-514) ----
      if (tid < (bp.foot_drift ? 4 * kH : 4)) {
        const int leg = tid & 3, step = tid >> 2;
        double R[9], I[9], T[9], Iw[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { R[i] = (double)st[kOffRot + i]; I[i] = bp.inertia[i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += R[3 * i + k] * I[3 * k + j];
            T[3 * i + j] = s;
          }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += T[3 * i + k] * R[3 * j + k];
            Iw[3 * i + j] = s;
          }
        // cofactor inverse (Eigen's fixed 3x3 inverse())
        const double c00 = Iw[4] * Iw[8] - Iw[5] * Iw[7];
        const double c01 = Iw[5] * Iw[6] - Iw[3] * Iw[8];
        const double c02 = Iw[3] * Iw[7] - Iw[4] * Iw[6];
        const double id = 1.0 / (Iw[0] * c00 + Iw[1] * c01 + Iw[2] * c02);
        double Inv[9];
        Inv[0] = c00 * id;
        Inv[1] = (Iw[2] * Iw[7] - Iw[1] * Iw[8]) * id;
        Inv[2] = (Iw[1] * Iw[5] - Iw[2] * Iw[4]) * id;
        Inv[3] = c01 * id;
        Inv[4] = (Iw[0] * Iw[8] - Iw[2] * Iw[6]) * id;
        Inv[5] = (Iw[2] * Iw[3] - Iw[0] * Iw[5]) * id;
        Inv[6] = c02 * id;
        Inv[7] = (Iw[1] * Iw[6] - Iw[0] * Iw[7]) * id;
        Inv[8] = (Iw[0] * Iw[4] - Iw[1] * Iw[3]) * id;
        double fx = st[kOffFoot + 3 * leg], fy = st[kOffFoot + 3 * leg + 1], fz = st[kOffFoot + 3 * leg + 2];
        if (bp.foot_drift) {
          // the body moves on with the commanded world velocity, the stance feet stay: r_i = r_0 - i dt v_d
          const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
          const double k = (double)step * bp.dt;
          fx -= k * (R[0] * vx + R[1] * vy + R[2] * vz);
          fy -= k * (R[3] * vx + R[4] * vy + R[5] * vz);
          fz -= k * (R[6] * vx + R[7] * vy + R[8] * vz);
        }
        // skew (Utils.cpp:35-41)
        const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += Inv[3 * i + k] * sk[3 * k + j];
            sm.Bd[step * 156 + (6 + i) * 12 + 3 * leg + j] = s * bp.dt;
            sm.Bd[step * 156 + (9 + i) * 12 + 3 * leg + j] = (i == j) ? (1.0 / bp.mass) * bp.dt : 0.0;
          }
        if (bp.exact_discretization) {
          // B_d += dt^2/2 A_c B_c: euler rows <- Rz' (I^-1 [r]x), position rows <- I / m
          const double yaw = (double)st[kOffEuler + 2];
          double sy, cy;
          sincos(yaw, &sy, &cy);
          const double h = 0.5 * bp.dt;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double b6 = sm.Bd[step * 156 + 6 * 12 + 3 * leg + j], b7 = sm.Bd[step * 156 + 7 * 12 + 3 * leg + j],
                         b8 = sm.Bd[step * 156 + 8 * 12 + 3 * leg + j];
            sm.Bd[step * 156 + 0 * 12 + 3 * leg + j] = h * (cy * b6 + sy * b7);
            sm.Bd[step * 156 + 1 * 12 + 3 * leg + j] = h * (-sy * b6 + cy * b7);
            sm.Bd[step * 156 + 2 * 12 + 3 * leg + j] = h * b8;
#pragma unroll
            for (int i = 0; i < 3; ++i)
              sm.Bd[step * 156 + (3 + i) * 12 + 3 * leg + j] = h * sm.Bd[step * 156 + (9 + i) * 12 + 3 * leg + j];
          }
        }
      }
      __syncthreads();
      // same B_d for every step unless the feet drift (A1RobotControl.cpp:498-514)
      if (!bp.foot_drift)
        for (int idx = tid; idx < (kH - 1) * 156; idx += kThreads) sm.Bd[156 + idx] = sm.Bd[idx % 156];
    } else {
      // ---- ConvexMpc surface: caller-written A_mat_d / B_mat_d_list ----
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = model.A_d[size_t(p) * 169 + tid];
      }
      for (int idx = tid; idx < kH * 156; idx += kThreads)
        sm.Bd[idx] = model.B_d_list[size_t(p) * kH * 156 + idx];
      if (tid < 13) sm.x0[tid] = model.x0[size_t(p) * 13 + tid];
      if (tid < kS) sm.xref[tid] = model.x_ref[size_t(p) * kS + tid];
      if (tid < 4 * kH) sm.contacts[tid] = model.contacts[size_t(p) * 4 + (tid & 3)] != 0;
    }
    // zero B_qp (upper blocks stay zero, ConvexMpc.cpp:94)
    for (int idx = tid; idx < kS * kNP / 2; idx += kThreads)
      reinterpret_cast<double2*>(sm.Bq)[idx] = make_double2(0.0, 0.0);
    __syncthreads();
    // ---- A_qp powers: block i = block (i-1) * A_d (ConvexMpc.cpp:185-191) ----
    for (int i = 1; i < kH; ++i) {
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        const double* Ap = &sm.Apow[i * 169];
        const double* A1 = &sm.Apow[169];
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 13; ++k) s += Ap[rr * 13 + k] * A1[k * 13 + cc];
        sm.Apow[(i + 1) * 169 + tid] = s;
      }
      __syncthreads();
    }
    // ---- B_qp block (i,j), j <= i: A_d^(i-j) * B_d(j) (ConvexMpc.cpp:192-201) ----
    for (int idx = tid; idx < (kH * (kH + 1) / 2) * 156; idx += kThreads) {
      const int blk = idx / 156, e = idx % 156;
      // blk -> (i, j) with j <= i, row-major over the lower triangle
      int i = 0, rem = blk;
      while (rem > i) { rem -= (i + 1); ++i; }
      const int j = rem;
      const int rr = e / 12, cc = e % 12;
      const double* Ap = &sm.Apow[(i - j) * 169];
      const double* Bj = &sm.Bd[j * 156];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) s += Ap[rr * 13 + k] * Bj[k * 12 + cc];
      sm.Bq[(13 * i + rr) * kNP + 12 * j + cc] = s;
    }
    // tmp = Q (A_qp x0 - x_ref) (ConvexMpc.cpp:215-216)
    if (tid < kS) {
      const int i = tid / 13, rr = tid % 13;
      const double* Ai = &sm.Apow[(i + 1) * 169 + rr * 13];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) s += Ai[k] * sm.x0[k];
      sm.tmp[tid] = bp.Qd[rr] * (s - sm.xref[tid]);
    }
    __syncthreads();
    // ---- K2: Hessian = B_qp' Q B_qp + R, fp64 accumulate (ConvexMpc.cpp:207-211) ----
    {
      double acc[4][8];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[rr][jj] = 0.0;
      // rows of B_qp above step block (4rg)/12 are structurally zero in these columns of B_qp'
      const int kstart = 13 * (rg / 3);
      for (int k = kstart; k < kS; ++k) {
        const double qk = bp.Qd[k % 13];
        const double2* rowp = reinterpret_cast<const double2*>(&sm.Bq[k * kNP + 4 * rg]);
        const double2 r01 = rowp[0], r23 = rowp[1];
        const double rop[4] = {r01.x * qk, r01.y * qk, r23.x * qk, r23.y * qk};
        const double2* colp = reinterpret_cast<const double2*>(&sm.Bq[k * kNP + 2 * cg]);
        double cop[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 v = colp[16 * i];
          cop[2 * i] = v.x;
          cop[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[rr][jj] = fma(rop[rr], cop[jj], acc[rr][jj]);
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int row = 4 * rg + rr;
        // P is handed to the solver in f64, rows padded to 128 (pad columns are exact zeros),
        // so one cp.async.bulk moves a whole problem into shared memory
        double* Pp = P_out + size_t(p) * kN * kNP + row * kNP;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = 32 * i + 2 * cg;
          double v0 = acc[rr][2 * i], v1 = acc[rr][2 * i + 1];
          if (c0 == row) v0 += bp.Rd[row % 12];
          if (c0 + 1 == row) v1 += bp.Rd[row % 12];
          *reinterpret_cast<double2*>(Pp + c0) = make_double2(v0, v1);
        }
      }
    }
    // ---- gradient = B_qp' tmp (ConvexMpc.cpp:217) ----
    if (tid < kN) {
      double s = 0.0;
      for (int k = 13 * (tid / 12); k < kS; ++k) s = fma(sm.Bq[k * kNP + tid], sm.tmp[k], s);
      q_out[size_t(p) * kN + tid] = s;
    }
    // ---- bounds, contacts replicated over the horizon (ConvexMpc.cpp:223-245) ----
    if (tid >= 256 && tid < 256 + kM) {
      const int i = tid - 256;
      const int leg = (i % 20) / 5, t = i % 5;
      const float cflag = sm.contacts[4 * (i / 20) + leg] ? 1.0f : 0.0f;
      float lo, hi;
      if (t == 0 || t == 2) { lo = 0.0f; hi = (float)MPC_INFTY; }
      else if (t == 1 || t == 3) { lo = -(float)MPC_INFTY; hi = 0.0f; }
      else { lo = (float)bp.fz_min * cflag; hi = (float)bp.fz_max * cflag; }
      l_out[size_t(p) * kM + i] = lo;
      u_out[size_t(p) * kM + i] = hi;
    }
  }
}

}  // namespace mpcb200
