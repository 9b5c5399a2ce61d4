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
// for STORAGE of P, q, l, u.  B200 runs DFMA at half the FFMA rate (measured
// 17.1 T DFMA/s), so the path computes in f64 and stores QP data in f32.
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
  int contacts[4];
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
qp_build_kernel(const MpcStateIn* __restrict__ states, ModelIn model, int num, float* __restrict__ P_out,
                float* __restrict__ q_out, float* __restrict__ l_out, float* __restrict__ u_out,
                const __grid_constant__ BuildParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
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
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = ((rr == cc) ? 1.0 : 0.0) + a * bp.dt;
      } else if (tid >= 192 && tid < 192 + 156) {
        sm.Bd[tid - 192] = 0.0;
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
      } else if (tid >= 416 && tid < 420) {
        sm.contacts[tid - 416] = st[kOffContacts + tid - 416] != 0.0f;
      }
      __syncthreads();
      // ---- B_d = dt*B_c, one thread per leg (ConvexMpc.cpp:132-143, :151) ----
      if (tid < 4) {
        const int leg = tid;
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
        const double fx = st[kOffFoot + 3 * leg], fy = st[kOffFoot + 3 * leg + 1],
                     fz = st[kOffFoot + 3 * leg + 2];
        // skew (Utils.cpp:35-41)
        const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += Inv[3 * i + k] * sk[3 * k + j];
            sm.Bd[(6 + i) * 12 + 3 * leg + j] = s * bp.dt;
            sm.Bd[(9 + i) * 12 + 3 * leg + j] = (i == j) ? (1.0 / bp.mass) * bp.dt : 0.0;
          }
      }
      __syncthreads();
      // same B_d for every step (A1RobotControl.cpp:498-514)
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
      if (tid < 4) sm.contacts[tid] = model.contacts[size_t(p) * 4 + tid] != 0;
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
        float* Pp = P_out + size_t(p) * kN * kN + row * kN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = 32 * i + 2 * cg;
          if (c0 < kN) {
            double v0 = acc[rr][2 * i], v1 = acc[rr][2 * i + 1];
            if (c0 == row) v0 += bp.Rd[row % 12];
            if (c0 + 1 == row) v1 += bp.Rd[row % 12];
            *reinterpret_cast<float2*>(Pp + c0) = make_float2((float)v0, (float)v1);
          }
        }
      }
    }
    // ---- gradient = B_qp' tmp (ConvexMpc.cpp:217) ----
    if (tid < kN) {
      double s = 0.0;
      for (int k = 13 * (tid / 12); k < kS; ++k) s = fma(sm.Bq[k * kNP + tid], sm.tmp[k], s);
      q_out[size_t(p) * kN + tid] = (float)s;
    }
    // ---- bounds, contacts replicated over the horizon (ConvexMpc.cpp:223-245) ----
    if (tid >= 256 && tid < 256 + kM) {
      const int i = tid - 256;
      const int leg = (i % 20) / 5, t = i % 5;
      const float cflag = sm.contacts[leg] ? 1.0f : 0.0f;
      float lo, hi;
      if (t == 0 || t == 2) { lo = 0.0f; hi = (float)MPC_INFTY; }
      else if (t == 1 || t == 3) { lo = -(float)MPC_INFTY; hi = 0.0f; }
      else { lo = (float)bp.fz_min * cflag; hi = (float)bp.fz_max * cflag; }
      l_out[size_t(p) * kM + i] = lo;
      u_out[size_t(p) * kM + i] = hi;
    }
  }
}

// ---------------------------------------------------------------------------
// K3+K4+K5: ADMM solve.  One CTA per problem, problems pulled from an atomic
// counter (iteration counts vary 100..400, so static assignment leaves a tail).
// ---------------------------------------------------------------------------
struct SolveSmem {
  double P[kN * kNP];       // unscaled Hessian as f64, row stride 128, cols 120..127 zero (122,880 B)
  double rhs[kNP];          // operand of the K^-1 matvec (pad = 0)
  double xD[kNP];           // D .* x for P x (pad = 0)
  double Dp[kNP];           // D (pad = 0)
  double buf[2][kNP];       // sweep: published pivot row (pad = 0)
  double piv[2][2];
  double x[kN], xt[kN], qb[kN], Dinv[kN], q0[kN];
  double z[kM], y[kM], lb[kM], ub[kM], E[kM], Einv[kM], rv[kM], rinv[kM], w[kM];
  double Av[kLegSteps * 9];  // scaled constraint entries per leg-step
  double G[kLegSteps * 9];   // A' diag(rho) A, 3x3 block per leg-step
  double red[kWarps * 16];
  double scal[8];            // 0:c 1:cinv 2:rho 3:ct 4:pri_res
  int flags[8];              // 0:done 1:status 2:refactor 3:problem index
  int ctype[kM];
};

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double limit_scaling(double v) {  // osqp scaling.c
  v = v < 1e-4 ? 1.0 : v;
  return v > 1e4 ? 1e4 : v;
}

// max_j |P_rj| * D_j for the thread's owned row (valid on the 4 lanes sharing it)
__device__ __forceinline__ double row_norm_pass(const SolveSmem& sm, int rg, int cg) {
  double dcol[8];
  const double2* dp = reinterpret_cast<const double2*>(&sm.Dp[2 * cg]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 v = dp[16 * i];
    dcol[2 * i] = v.x;
    dcol[2 * i + 1] = v.y;
  }
  double m[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const double2* pp = reinterpret_cast<const double2*>(&sm.P[(4 * rg + rr) * kNP + 2 * cg]);
    double mm = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 v = pp[16 * i];
      mm = fmax(mm, fabs(v.x) * dcol[2 * i]);
      mm = fmax(mm, fabs(v.y) * dcol[2 * i + 1]);
    }
    m[rr] = mm;
  }
  return reduce_scatter_max(m, cg);
}

// Build K = c D P D + sigma I + A' diag(rho) A into the register tiles, then
// overwrite it with -K^-1 by the symmetric sweep operator (one pivot per step).
// Step k: the 16 threads of row group k/4 publish row k (the pivot d replaced
// by d-1 so the generic rank-1 update also produces column k); everybody applies
//   a_rj <- a_rj - (a_kr / d) * a'_kj      (r != k)
//   a_kj <- a_kj / d,  a_kk <- -1/d        (r == k)
__device__ __forceinline__ void factor_inverse(SolveSmem& sm, double (&a)[4][8], int rg, int cg,
                                               int tid, double sigma) {
  if (tid < kLegSteps * 9) {
    const int k = tid / 9, rr = (tid % 9) / 3, cc = tid % 3;
    const double* av = &sm.Av[k * 9];
    // row e coefficients on (x, y, z): e0:(av0,0,av1) e1:(av2,0,av3) e2:(0,av4,av5) e3:(0,av6,av7) e4:(0,0,av8)
    double g = 0.0;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      double co[3];
      co[0] = (e == 0) ? av[0] : (e == 1) ? av[2] : 0.0;
      co[1] = (e == 2) ? av[4] : (e == 3) ? av[6] : 0.0;
      co[2] = (e == 0) ? av[1] : (e == 1) ? av[3] : (e == 2) ? av[5] : (e == 3) ? av[7] : av[8];
      g += sm.rv[5 * k + e] * co[rr] * co[cc];
    }
    sm.G[tid] = g;
  }
  __syncthreads();
  {
    const double c = sm.scal[0];
    double dcol[8];
    const double2* dp = reinterpret_cast<const double2*>(&sm.Dp[2 * cg]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 v = dp[16 * i];
      dcol[2 * i] = v.x;
      dcol[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int row = 4 * rg + rr;
      const double cDr = c * sm.Dp[row];
      const double2* pp = reinterpret_cast<const double2*>(&sm.P[row * kNP + 2 * cg]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 v = pp[16 * i];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 32 * i + 2 * cg + e;
          double val = cDr * (e ? v.y : v.x) * dcol[2 * i + e];
          if (col == row) val += sigma;
          if (col / 3 == row / 3) val += sm.G[(row / 3) * 9 + (row % 3) * 3 + (col % 3)];
          a[rr][2 * i + e] = val;
        }
      }
    }
  }
  for (int k = 0; k < kN; ++k) {
    const int cur = k & 1;
    const bool mine = (rg == (k >> 2));
    const int kr = k & 3;
    if (mine) {
      // publish row k; its owner of column k swaps the pivot d for d-1
      double v[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        v[jj] = (kr == 0) ? a[0][jj] : (kr == 1) ? a[1][jj] : (kr == 2) ? a[2][jj] : a[3][jj];
      if (((k >> 1) & 15) == cg) {
        const int kj = 2 * (k >> 5) + (k & 1);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (jj == kj) { sm.piv[cur][0] = v[jj]; v[jj] -= 1.0; }
      }
      double2* dst = reinterpret_cast<double2*>(&sm.buf[cur][2 * cg]);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[16 * i] = make_double2(v[2 * i], v[2 * i + 1]);
    }
    __syncthreads();
    const double dinv = __drcp_rn(sm.piv[cur][0]);
    double vcol[8];
    {
      const double2* src = reinterpret_cast<const double2*>(&sm.buf[cur][2 * cg]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 v = src[16 * i];
        vcol[2 * i] = v.x;
        vcol[2 * i + 1] = v.y;
      }
    }
    // a_rk == a_kr by symmetry, so the published row also supplies column k
    const double2* rsrc = reinterpret_cast<const double2*>(&sm.buf[cur][4 * rg]);
    const double2 w01 = rsrc[0], w23 = rsrc[1];
    const double w[4] = {-w01.x * dinv, -w01.y * dinv, -w23.x * dinv, -w23.y * dinv};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      if (mine && rr == kr) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int col = 32 * (jj >> 1) + 2 * cg + (jj & 1);
          a[rr][jj] = (col == k) ? -dinv : a[rr][jj] * dinv;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) a[rr][jj] = fma(w[rr], vcol[jj], a[rr][jj]);
      }
    }
  }
  __syncthreads();
}

// x~ = K^-1 rhs with a = -K^-1 in the register tiles; returns x~ of the thread's owned row
__device__ __forceinline__ double kinv_matvec(const SolveSmem& sm, const double (&a)[4][8], int cg) {
  const double2* src = reinterpret_cast<const double2*>(&sm.rhs[2 * cg]);
  double v[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 t = src[16 * i];
    v[2 * i] = t.x;
    v[2 * i + 1] = t.y;
  }
  double s[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    double s0 = a[rr][0] * v[0], s1 = a[rr][1] * v[1];
#pragma unroll
    for (int jj = 2; jj < 8; jj += 2) {
      s0 = fma(a[rr][jj], v[jj], s0);
      s1 = fma(a[rr][jj + 1], v[jj + 1], s1);
    }
    s[rr] = s0 + s1;
  }
  return -reduce_scatter_sum(s, cg);
}

// registers are allocated per 4 warps: 480 threads count as 16 warps, so 128 is the cap
// (__maxnreg__(136) fails to launch: "too many resources requested")
__global__ void __launch_bounds__(kThreads, 1)
admm_solve_kernel(const float* __restrict__ P_all, const float* __restrict__ q_all,
                  const float* __restrict__ l_all, const float* __restrict__ u_all,
                  const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                  float* __restrict__ x_all, int num, int* __restrict__ counter,
                  const __grid_constant__ SolveParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SolveSmem& sm = *reinterpret_cast<SolveSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = tid >> 4, cg = tid & 15;
  // variable ownership: lanes with cg % 4 == 0 own row r of their row group
  const bool vown = (cg & 3) == 0;
  const int r = 4 * rg + owned_row(cg);
  // constraint-row ownership: 6 leg-steps (30 lanes) per warp, warps 0..6
  const int ck = warp * 6 + lane / 5;  // leg-step
  const int ce = lane % 5;             // row inside the leg-step
  const bool crow = (warp < 7) && (lane < 30) && (ck < kLegSteps);
  const int ci = crow ? 5 * ck + ce : 0;                   // constraint row
  const int cja = crow ? 3 * ck + ((ce < 2) ? 0 : 1) : 0;  // lateral variable of the row
  const int cjz = crow ? 3 * ck + 2 : 0;
  const double mu = sp.mu;

  double a[4][8];  // register tile of -K^-1

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;

    // ---- load P (coalesced float4 -> f64 smem), q, l, u ----
    {
      const float4* src = reinterpret_cast<const float4*>(P_all + size_t(p) * kN * kN);
      for (int idx = tid; idx < kN * kN / 4; idx += kThreads) {
        const float4 v = __ldg(src + idx);
        const int e = idx * 4, rr = e / kN, cc = e % kN;  // 120 % 4 == 0: never straddles rows
        double2* dst = reinterpret_cast<double2*>(&sm.P[rr * kNP + cc]);
        dst[0] = make_double2((double)v.x, (double)v.y);
        dst[1] = make_double2((double)v.z, (double)v.w);
      }
      if (tid < kN) {
        double2* padp = reinterpret_cast<double2*>(&sm.P[tid * kNP + kN]);
#pragma unroll
        for (int i = 0; i < 4; ++i) padp[i] = make_double2(0.0, 0.0);
      }
      if (tid < kNP) {
        sm.Dp[tid] = (tid < kN) ? 1.0 : 0.0;
        sm.rhs[tid] = 0.0;
        sm.xD[tid] = 0.0;
        sm.buf[0][tid] = 0.0;
        sm.buf[1][tid] = 0.0;
      }
      if (tid < kN) {
        sm.q0[tid] = (double)q_all[size_t(p) * kN + tid];
        sm.x[tid] = 0.0;
        sm.xt[tid] = 0.0;
      }
      if (tid >= 256 && tid < 256 + kM) {
        const int i = tid - 256;
        sm.lb[i] = (double)l_all[size_t(p) * kM + i];
        sm.ub[i] = (double)u_all[size_t(p) * kM + i];
        sm.E[i] = 1.0;
        sm.z[i] = 0.0;
        sm.y[i] = 0.0;
      }
      if (tid == 0) {
        sm.scal[0] = 1.0;
        sm.scal[2] = sp.rho;
        sm.flags[0] = 0;
        sm.flags[1] = MPC_STATUS_UNSOLVED;
      }
    }
    __syncthreads();

    // ---- K3a: modified Ruiz equilibration (osqp scaling.c scale_data) ----
    // scaled quantities are never materialised: P_bar = c D P D, A_bar = E A D.
    if (sp.scaling > 0) {
      double nP = row_norm_pass(sm, rg, cg);  // c = 1, D = 1
      for (int it = 0; it < sp.scaling; ++it) {
        double Dt = 1.0, Et = 1.0;
        if (vown) {
          // column norm of [P; A] for variable r
          const int k = r / 3, c3 = r % 3;
          const double* Ek = &sm.E[5 * k];
          double nA;
          if (c3 == 0) nA = fmax(Ek[0], Ek[1]);
          else if (c3 == 1) nA = fmax(Ek[2], Ek[3]);
          else nA = fmax(mu * fmax(fmax(Ek[0], Ek[1]), fmax(Ek[2], Ek[3])), Ek[4]);
          nA *= sm.Dp[r];
          Dt = rsqrt(limit_scaling(fmax(nP, nA)));
        }
        if (crow) {
          // row norm of A for constraint ci
          const double dz = sm.Dp[cjz];
          const double nrow = (ce == 4) ? dz : fmax(sm.Dp[cja], mu * dz);
          Et = rsqrt(limit_scaling(sm.E[ci] * nrow));
        }
        __syncthreads();
        if (vown) sm.Dp[r] *= Dt;
        if (crow) sm.E[ci] *= Et;
        __syncthreads();
        // cost normalisation with the new D and the old c
        const double c_old = sm.scal[0];
        const double nP2 = c_old * sm.Dp[r] * row_norm_pass(sm, rg, cg);
        double part_sum = vown ? nP2 : 0.0;
        double part_q = vown ? fabs(c_old * sm.Dp[r] * sm.q0[r]) : 0.0;
        part_sum = warp_sum(part_sum);
        part_q = warp_max(part_q);
        if (lane == 0) {
          sm.red[warp * 16 + 0] = part_sum;
          sm.red[warp * 16 + 1] = part_q;
        }
        __syncthreads();
        if (tid == 0) {
          double s = 0.0, qn = 0.0;
          for (int w = 0; w < kWarps; ++w) {
            s += sm.red[w * 16 + 0];
            qn = fmax(qn, sm.red[w * 16 + 1]);
          }
          const double mean = s / (double)kN;
          const double ct = 1.0 / limit_scaling(fmax(mean, limit_scaling(qn)));
          sm.scal[3] = ct;
          sm.scal[0] = c_old * ct;
        }
        __syncthreads();
        nP = nP2 * sm.scal[3];
      }
    }
    // ---- scaled data: q_bar, bounds, constraint entries, rho vector ----
    {
      const double c = sm.scal[0];
      if (tid == 0) sm.scal[1] = 1.0 / c;
      if (tid < kN) {
        const double d = sm.Dp[tid];
        sm.qb[tid] = c * d * sm.q0[tid];
        sm.Dinv[tid] = 1.0 / d;
      }
      if (crow) {
        const double e = sm.E[ci];
        const double l = e * sm.lb[ci], u = e * sm.ub[ci];
        sm.lb[ci] = l;
        sm.ub[ci] = u;
        sm.Einv[ci] = 1.0 / e;
        int ct = 0;
        if (l < -MPC_INFTY * 1e-4 && u > MPC_INFTY * 1e-4) ct = -1;
        else if (u - l < 1e-4) ct = 1;
        sm.ctype[ci] = ct;
        const double rho = sp.rho;
        const double rvv = (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho;
        sm.rv[ci] = rvv;
        sm.rinv[ci] = 1.0 / rvv;
      }
      if (tid >= 256 && tid < 256 + kLegSteps) {
        const int k = tid - 256;
        const double dx = sm.Dp[3 * k], dy = sm.Dp[3 * k + 1], dz = sm.Dp[3 * k + 2];
        const double* e = &sm.E[5 * k];
        double* av = &sm.Av[9 * k];
        av[0] = e[0] * dx;  av[1] = mu * e[0] * dz;
        av[2] = e[1] * dx;  av[3] = -mu * e[1] * dz;
        av[4] = e[2] * dy;  av[5] = mu * e[2] * dz;
        av[6] = e[3] * dy;  av[7] = -mu * e[3] * dz;
        av[8] = e[4] * dz;
      }
    }
    __syncthreads();
    // per-row constraint coefficients (registers)
    double cca = 0.0, ccz = 0.0;
    if (crow) {
      const double* av = &sm.Av[9 * ck];
      if (ce < 4) { cca = av[2 * ce]; ccz = av[2 * ce + 1]; }
      else { cca = 0.0; ccz = av[8]; }
    }
    // first rhs: x = z = y = 0  ->  rhs = -q_bar
    if (tid < kN) sm.rhs[tid] = -sm.qb[tid];

    // ---- K3b: factor (explicit inverse in registers) ----
    factor_inverse(sm, a, rg, cg, tid, sp.sigma);

    // ---- K4: ADMM iterations (osqp.c osqp_solve) ----
    const double sigma = sp.sigma, alpha = sp.alpha;
    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    for (iter = 1; iter <= sp.max_iter; ++iter) {
      // x~ = K^-1 rhs ; x <- alpha x~ + (1-alpha) x
      const double xt = kinv_matvec(sm, a, cg);
      if (vown) {
        sm.xt[r] = xt;
        sm.x[r] = alpha * xt + (1.0 - alpha) * sm.x[r];
      }
      __syncthreads();
      // z~ = A x~ ; z, y update ; next rhs = sigma x - q + A'(rho z - y)
      if (warp < 7) {
        if (crow) {
          const double zt = cca * sm.xt[cja] + ccz * sm.xt[cjz];
          const double zr = alpha * zt + (1.0 - alpha) * sm.z[ci];
          const double rvv = sm.rv[ci];
          const double yo = sm.y[ci];
          double zn = zr + sm.rinv[ci] * yo;
          zn = fmin(fmax(zn, sm.lb[ci]), sm.ub[ci]);
          const double yn = yo + rvv * (zr - zn);
          sm.z[ci] = zn;
          sm.y[ci] = yn;
          sm.w[ci] = rvv * zn - yn;
        }
        __syncwarp();
        if (crow && ce < 3) {
          const int j = 3 * ck + ce;
          const double* av = &sm.Av[9 * ck];
          const double* w = &sm.w[5 * ck];
          double s;
          if (ce == 0) s = av[0] * w[0] + av[2] * w[1];
          else if (ce == 1) s = av[4] * w[2] + av[6] * w[3];
          else s = av[1] * w[0] + av[3] * w[1] + av[5] * w[2] + av[7] * w[3] + av[8] * w[4];
          sm.rhs[j] = sigma * sm.x[j] - sm.qb[j] + s;
        }
      }
      const bool can_check = sp.check_termination > 0 && (iter % sp.check_termination == 0);
      const bool can_adapt = sp.adaptive_rho && sp.adaptive_rho_interval > 0 &&
                             (iter % sp.adaptive_rho_interval == 0);
      const bool last = (iter == sp.max_iter);
      if (!(can_check || can_adapt || last)) {
        __syncthreads();
        continue;
      }
      // ---- residuals (auxil.c compute_pri_res / compute_dua_res / tolerances) ----
      if (tid < kN) sm.xD[tid] = sm.Dp[tid] * sm.x[tid];
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      if (crow) {
        const double Ax = cca * sm.x[cja] + ccz * sm.x[cjz];
        const double zz = sm.z[ci];
        const double rp_ = Ax - zz;
        const double ei = sm.Einv[ci];
        v[0] = fabs(rp_);        // scaled primal residual
        v[1] = fabs(ei * rp_);   // unscaled
        v[2] = fabs(ei * zz);
        v[3] = fabs(ei * Ax);
        v[4] = fabs(zz);
        v[5] = fabs(Ax);
      }
      {
        // P_bar x = c D (P (D x))
        double xv[8];
        const double2* xp = reinterpret_cast<const double2*>(&sm.xD[2 * cg]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 t = xp[16 * i];
          xv[2 * i] = t.x;
          xv[2 * i + 1] = t.y;
        }
        double s[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const double2* pp = reinterpret_cast<const double2*>(&sm.P[(4 * rg + rr) * kNP + 2 * cg]);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double2 t = pp[16 * i];
            s0 = fma(t.x, xv[2 * i], s0);
            s1 = fma(t.y, xv[2 * i + 1], s1);
          }
          s[rr] = s0 + s1;
        }
        const double sr = reduce_scatter_sum(s, cg);
        if (vown) {
          const double Px = sm.scal[0] * sm.Dp[r] * sr;
          const int k = r / 3, c3 = r % 3;
          const double* av = &sm.Av[9 * k];
          const double* yy = &sm.y[5 * k];
          double Aty;
          if (c3 == 0) Aty = av[0] * yy[0] + av[2] * yy[1];
          else if (c3 == 1) Aty = av[4] * yy[2] + av[6] * yy[3];
          else Aty = av[1] * yy[0] + av[3] * yy[1] + av[5] * yy[2] + av[7] * yy[3] + av[8] * yy[4];
          const double qq = sm.qb[r];
          const double rd = Px + qq + Aty;
          const double di = sm.Dinv[r];
          v[6] = fabs(rd);        // scaled dual residual
          v[7] = fabs(di * rd);   // unscaled (times cinv later)
          v[8] = fmax(fmax(fabs(di * qq), fabs(di * Aty)), fabs(di * Px));
          v[9] = fmax(fmax(fabs(qq), fabs(Aty)), fabs(Px));
        }
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const double m = warp_max(v[i]);
        if (lane == 0) sm.red[warp * 16 + i] = m;
      }
      __syncthreads();
      if (tid == 0) {
        double m[10];
        for (int i = 0; i < 10; ++i) {
          double t = 0.0;
          for (int w = 0; w < kWarps; ++w) t = fmax(t, sm.red[w * 16 + i]);
          m[i] = t;
        }
        const double cinv = sm.scal[1];
        const double pri = m[1], dua = cinv * m[7];
        const double eps_pri = sp.eps_abs + sp.eps_rel * fmax(m[2], m[3]);
        const double eps_dua = sp.eps_abs + sp.eps_rel * cinv * m[8];
        sm.scal[4] = pri;
        int done = 0, refactor = 0;
        if ((can_check || last) && pri < eps_pri && dua < eps_dua) {
          done = 1;
          sm.flags[1] = MPC_STATUS_SOLVED;
        } else if (last) {
          // osqp.c: approximate check at 10x tolerances, else MAX_ITER_REACHED
          done = 1;
          sm.flags[1] = (pri < 10.0 * eps_pri && dua < 10.0 * eps_dua) ? 2 : MPC_STATUS_MAX_ITER_REACHED;
        } else if (can_adapt) {
          // auxil.c compute_rho_estimate / adapt_rho (scaled quantities)
          const double rho = sm.scal[2];
          const double pn = m[0] / (fmax(m[4], m[5]) + 1e-10);
          const double dn = m[6] / (m[9] + 1e-10);
          double rho_new = rho * sqrt(pn / (dn + 1e-10));
          rho_new = fmin(fmax(rho_new, 1e-6), 1e6);
          if (rho_new > rho * sp.adaptive_rho_tolerance || rho_new < rho / sp.adaptive_rho_tolerance) {
            sm.scal[2] = rho_new;
            refactor = 1;
          }
        }
        sm.flags[0] = done;
        sm.flags[2] = refactor;
      }
      __syncthreads();
      if (sm.flags[0]) {
        status = sm.flags[1];
        pri_res_out = sm.scal[4];
        break;
      }
      if (sm.flags[2]) {
        ++rho_updates;
        const double rho = sm.scal[2];
        if (crow) {
          const int ct = sm.ctype[ci];
          const double rvv = (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho;
          sm.rv[ci] = rvv;
          sm.rinv[ci] = 1.0 / rvv;
          // rhs was built with the old rho vector: rebuild it
          sm.w[ci] = rvv * sm.z[ci] - sm.y[ci];
        }
        __syncwarp();
        if (crow && ce < 3) {
          const int j = 3 * ck + ce;
          const double* av = &sm.Av[9 * ck];
          const double* w = &sm.w[5 * ck];
          double s;
          if (ce == 0) s = av[0] * w[0] + av[2] * w[1];
          else if (ce == 1) s = av[4] * w[2] + av[6] * w[3];
          else s = av[1] * w[0] + av[3] * w[1] + av[5] * w[2] + av[7] * w[3] + av[8] * w[4];
          sm.rhs[j] = sigma * sm.x[j] - sm.qb[j] + s;
        }
        __syncthreads();
        factor_inverse(sm, a, rg, cg, tid, sigma);
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    __syncthreads();
    if (x_all != nullptr && tid < kN) x_all[size_t(p) * kN + tid] = (float)(sm.Dp[tid] * sm.x[tid]);
    if (tid < 12) {
      const int leg = tid / 3, rr = tid % 3;
      const double f0 = sm.Dp[3 * leg] * sm.x[3 * leg];
      const double f1 = sm.Dp[3 * leg + 1] * sm.x[3 * leg + 1];
      const double f2 = sm.Dp[3 * leg + 2] * sm.x[3 * leg + 2];
      double g;
      if (states != nullptr) {
        // R' f (A1RobotControl.cpp:558-561)
        const float* R = reinterpret_cast<const float*>(states + p) + kOffRot;
        g = (double)R[rr] * f0 + (double)R[3 + rr] * f1 + (double)R[6 + rr] * f2;
      } else {
        g = (rr == 0) ? f0 : (rr == 1) ? f1 : f2;
      }
      const bool bad = isnan(f0) || isnan(f1) || isnan(f2);  // NaN guard (:559)
      results[p].grf[tid] = bad ? 0.0f : (float)g;
    }
    if (tid == 32) {
      results[p].status = status;
      results[p].iters = iter;
      results[p].rho_updates = rho_updates;
      results[p].pri_res = (float)pri_res_out;
    }
  }
}

}  // namespace mpcb200
