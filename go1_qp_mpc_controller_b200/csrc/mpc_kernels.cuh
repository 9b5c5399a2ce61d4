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
// for STORAGE of P, q, l, u and for the Ruiz norm passes.  B200 runs DFMA at
// half the FFMA rate, so the path computes in f64 and stores QP data in f32.
//
// Thread layout shared by both kernels (H = 10, n = 120): 480 threads =
// 120 rows x 4 parts; thread (r, part) owns columns part*30 .. part*30+29 of
// row r.  Rows of operand vectors / matrices are stored "chunk padded": chunk
// `part` starts at element part*32 so every thread's 30 operands are 16 B
// aligned (LDS.128) and bank-conflict free (4 distinct addresses per warp).
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
constexpr int kChunk = 30;
constexpr int kChunkPad = 32;
constexpr int kRowPad = 128;
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

__host__ __device__ constexpr int padj(int j) { return (j / kChunk) * kChunkPad + (j % kChunk); }

// ---------------------------------------------------------------------------
// K0+K1+K2: QP build.  One CTA per problem, grid-stride over problems.
// ---------------------------------------------------------------------------
struct BuildSmem {
  double Bq[kS * kRowPad];       // B_qp, rows chunk-padded (133,120 B)
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
  const int r = tid >> 2, part = tid & 3;

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
    for (int idx = tid; idx < kS * kRowPad / 2; idx += kThreads)
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
      sm.Bq[(13 * i + rr) * kRowPad + padj(12 * j + cc)] = s;
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
      double acc[kChunk];
#pragma unroll
      for (int jj = 0; jj < kChunk; ++jj) acc[jj] = 0.0;
      const int rb = r / 12, cb0 = (part * kChunk) / 12;
      const int kstart = 13 * (rb > cb0 ? rb : cb0);  // rows above are structurally zero
      const int rp = padj(r);
      for (int k = kstart; k < kS; ++k) {
        const double b = sm.Bq[k * kRowPad + rp] * bp.Qd[k % 13];
        const double2* row = reinterpret_cast<const double2*>(&sm.Bq[k * kRowPad + part * kChunkPad]);
#pragma unroll
        for (int j2 = 0; j2 < kChunk / 2; ++j2) {
          const double2 v = row[j2];
          acc[2 * j2] = fma(b, v.x, acc[2 * j2]);
          acc[2 * j2 + 1] = fma(b, v.y, acc[2 * j2 + 1]);
        }
      }
      float* Pp = P_out + size_t(p) * kN * kN + r * kN + part * kChunk;
#pragma unroll
      for (int j2 = 0; j2 < kChunk / 2; ++j2) {
        const int c0 = part * kChunk + 2 * j2;
        double v0 = acc[2 * j2], v1 = acc[2 * j2 + 1];
        if (c0 == r) v0 += bp.Rd[r % 12];
        if (c0 + 1 == r) v1 += bp.Rd[r % 12];
        reinterpret_cast<float2*>(Pp)[j2] = make_float2((float)v0, (float)v1);
      }
    }
    // ---- gradient = B_qp' tmp (ConvexMpc.cpp:217) ----
    if (tid < kN) {
      const int cp = padj(tid);
      double s = 0.0;
      for (int k = 13 * (tid / 12); k < kS; ++k) s = fma(sm.Bq[k * kRowPad + cp], sm.tmp[k], s);
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
  float P[kN * kRowPad];    // unscaled Hessian, rows chunk-padded (61,440 B)
  double rhs[kRowPad];      // chunk-padded: operand of the K^-1 matvec
  double xD[kRowPad];       // chunk-padded: D .* x for P x
  double Dp[kRowPad];       // chunk-padded D
  double buf[2][kRowPad];   // sweep: published pivot row
  float Df[kRowPad];        // chunk-padded float copy of D for the norm passes
  double piv[2][2];
  double x[kN], xt[kN], qb[kN], D[kN], Dinv[kN], q0[kN];
  double z[kM], y[kM], lb[kM], ub[kM], E[kM], Einv[kM], rv[kM], rinv[kM], w[kM];
  double Av[kLegSteps * 9];  // scaled constraint entries per leg-step
  double G[kLegSteps * 9];   // A' diag(rho) A, 3x3 block per leg-step
  double red[kWarps * 16];
  double scal[8];            // 0:c 1:cinv 2:rho 3:ct
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

// max_j |P_rj| * D_j over the whole row r (fp32), valid on all 4 lanes of the row
__device__ __forceinline__ float row_norm_pass(const SolveSmem& sm, int r, int part) {
  const float4* prow = reinterpret_cast<const float4*>(&sm.P[r * kRowPad + part * kChunkPad]);
  const float4* drow = reinterpret_cast<const float4*>(&sm.Df[part * kChunkPad]);
  float m = 0.0f;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {  // pad lanes hold P = 0
    const float4 a = prow[j4], d = drow[j4];
    m = fmaxf(m, fabsf(a.x) * d.x);
    m = fmaxf(m, fabsf(a.y) * d.y);
    m = fmaxf(m, fabsf(a.z) * d.z);
    m = fmaxf(m, fabsf(a.w) * d.w);
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  return m;
}

// Build K = c D P D + sigma I + A' diag(rho) A into registers, then overwrite it
// with -K^-1 by the symmetric sweep operator (one pivot per step, n steps).
// Step k: the 4 threads of row k publish the row (with the pivot replaced by
// d-1 so the generic rank-1 update also produces column k), everybody applies
//   a_rj <- a_rj - (a_kr / d) * a'_kj      (r != k)
//   a_kj <- a_kj / d,  a_kk <- -1/d        (r == k)
__device__ __forceinline__ void factor_inverse(SolveSmem& sm, double (&a)[kChunk], int r, int part,
                                               int tid, double sigma) {
  // G blocks
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
    const double cDr = sm.scal[0] * sm.D[r];
    const float* prow = &sm.P[r * kRowPad + part * kChunkPad];
    const double* drow = &sm.Dp[part * kChunkPad];
#pragma unroll
    for (int jj = 0; jj < kChunk; ++jj) {
      const int col = part * kChunk + jj;
      double v = cDr * (double)prow[jj] * drow[jj];
      if (col == r) v += sigma;
      if (col / 3 == r / 3) v += sm.G[(r / 3) * 9 + (r % 3) * 3 + (col % 3)];
      a[jj] = v;
    }
  }
  const int rp = padj(r);
  for (int k = 0; k < kN; ++k) {
    const int cur = k & 1;
    if (r == k) {
      double2* dst = reinterpret_cast<double2*>(&sm.buf[cur][part * kChunkPad]);
      const bool has_pivot = (k / kChunk) == part;
      const int kj = k - part * kChunk;
#pragma unroll
      for (int j2 = 0; j2 < kChunk / 2; ++j2) {
        double v0 = a[2 * j2], v1 = a[2 * j2 + 1];
        if (has_pivot && kj == 2 * j2) { sm.piv[cur][0] = v0; v0 -= 1.0; }
        if (has_pivot && kj == 2 * j2 + 1) { sm.piv[cur][0] = v1; v1 -= 1.0; }
        dst[j2] = make_double2(v0, v1);
      }
    }
    __syncthreads();
    const double d = sm.piv[cur][0];
    const double dinv = __drcp_rn(d);
    const double2* src = reinterpret_cast<const double2*>(&sm.buf[cur][part * kChunkPad]);
    if (r != k) {
      // a_rk == a_kr by symmetry, so the published row also supplies column k;
      // buf[k] itself was replaced by d-1 (r != k never reads it as vr)
      const double vr = sm.buf[cur][rp];
      const double w = -vr * dinv;
#pragma unroll
      for (int j2 = 0; j2 < kChunk / 2; ++j2) {
        const double2 v = src[j2];
        a[2 * j2] = fma(w, v.x, a[2 * j2]);
        a[2 * j2 + 1] = fma(w, v.y, a[2 * j2 + 1]);
      }
    } else {
      const int kj = k - part * kChunk;  // only meaningful when this part holds the pivot
      const bool has_pivot = (k / kChunk) == part;
#pragma unroll
      for (int jj = 0; jj < kChunk; ++jj) {
        double v = a[jj] * dinv;
        if (has_pivot && jj == kj) v = -dinv;
        a[jj] = v;
      }
    }
  }
  __syncthreads();
}

// x~ = K^-1 rhs with a = -K^-1 in registers; result valid on all 4 lanes of row r
__device__ __forceinline__ double kinv_matvec(const SolveSmem& sm, const double (&a)[kChunk], int part) {
  const double2* src = reinterpret_cast<const double2*>(&sm.rhs[part * kChunkPad]);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int j2 = 0; j2 < kChunk / 2; j2 += 3) {
    const double2 v0 = src[j2], v1 = src[j2 + 1], v2 = src[j2 + 2];
    s0 = fma(a[2 * j2], v0.x, s0);
    s0 = fma(a[2 * j2 + 1], v0.y, s0);
    s1 = fma(a[2 * j2 + 2], v1.x, s1);
    s1 = fma(a[2 * j2 + 3], v1.y, s1);
    s2 = fma(a[2 * j2 + 4], v2.x, s2);
    s2 = fma(a[2 * j2 + 5], v2.y, s2);
  }
  double s = (s0 + s1) + s2;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return -s;
}

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
  const int r = tid >> 2, part = tid & 3;
  // constraint-row ownership: 6 leg-steps (30 lanes) per warp, warps 0..6
  const int ck = warp * 6 + lane / 5;  // leg-step
  const int ce = lane % 5;             // row inside the leg-step
  const bool crow = (warp < 7) && (lane < 30) && (ck < kLegSteps);
  const int ci = 5 * ck + ce;          // constraint row
  const int cja = 3 * ck + ((ce < 2) ? 0 : 1);  // lateral variable of the row
  const int cjz = 3 * ck + 2;
  const double mu = sp.mu;

  double a[kChunk];  // row r, columns part*30.. of -K^-1

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;

    // ---- load P (coalesced float4), q, l, u ----
    {
      const float4* src = reinterpret_cast<const float4*>(P_all + size_t(p) * kN * kN);
      for (int idx = tid; idx < kN * kN / 4; idx += kThreads) {
        const float4 v = src[idx];
        const int e = idx * 4, rr = e / kN, cc = e % kN;  // 120 % 4 == 0: a float4 never straddles rows
        float* dst = &sm.P[rr * kRowPad];
        dst[padj(cc)] = v.x;
        dst[padj(cc + 1)] = v.y;
        dst[padj(cc + 2)] = v.z;
        dst[padj(cc + 3)] = v.w;
      }
      // pad lanes of every chunk must read as zero in the norm passes
      if (tid < kN * 4) {
        sm.P[(tid >> 2) * kRowPad + (tid & 3) * kChunkPad + 30] = 0.0f;
        sm.P[(tid >> 2) * kRowPad + (tid & 3) * kChunkPad + 31] = 0.0f;
      }
      if (tid < kRowPad) {
        sm.Df[tid] = 1.0f;
        sm.Dp[tid] = 1.0;
        sm.rhs[tid] = 0.0;
        sm.xD[tid] = 0.0;
        sm.buf[0][tid] = 0.0;
        sm.buf[1][tid] = 0.0;
      }
      if (tid < kN) {
        sm.q0[tid] = (double)q_all[size_t(p) * kN + tid];
        sm.D[tid] = 1.0;
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
      double nP = (double)row_norm_pass(sm, r, part);  // c = 1, D = 1
      for (int it = 0; it < sp.scaling; ++it) {
        double Dt = 1.0, Et = 1.0;
        if (part == 0) {
          // column norm of [P; A] for variable r
          const int k = r / 3, c3 = r % 3;
          const double* Ek = &sm.E[5 * k];
          double nA;
          if (c3 == 0) nA = fmax(Ek[0], Ek[1]);
          else if (c3 == 1) nA = fmax(Ek[2], Ek[3]);
          else nA = fmax(mu * fmax(fmax(Ek[0], Ek[1]), fmax(Ek[2], Ek[3])), Ek[4]);
          nA *= sm.D[r];
          Dt = rsqrt(limit_scaling(fmax(nP, nA)));
        }
        if (crow) {
          // row norm of A for constraint ci
          const double dz = sm.D[cjz];
          const double nrow = (ce == 4) ? dz : fmax(sm.D[cja], mu * dz);
          Et = rsqrt(limit_scaling(sm.E[ci] * nrow));
        }
        __syncthreads();
        if (part == 0) {
          const double dn = sm.D[r] * Dt;
          sm.D[r] = dn;
          sm.Dp[padj(r)] = dn;
          sm.Df[padj(r)] = (float)dn;
        }
        if (crow) sm.E[ci] *= Et;
        __syncthreads();
        // cost normalisation with the new D and the old c
        const double c_old = sm.scal[0];
        const double nP2 = c_old * sm.D[r] * (double)row_norm_pass(sm, r, part);
        double part_sum = (part == 0) ? nP2 : 0.0;
        double part_q = (part == 0) ? fabs(c_old * sm.D[r] * sm.q0[r]) : 0.0;
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
        sm.qb[tid] = c * sm.D[tid] * sm.q0[tid];
        sm.Dinv[tid] = 1.0 / sm.D[tid];
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
        const double dx = sm.D[3 * k], dy = sm.D[3 * k + 1], dz = sm.D[3 * k + 2];
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
    if (tid < kN) sm.rhs[padj(tid)] = -sm.qb[tid];

    // ---- K3b: factor (explicit inverse in registers) ----
    factor_inverse(sm, a, r, part, tid, sp.sigma);

    // ---- K4: ADMM iterations (osqp.c osqp_solve) ----
    const double sigma = sp.sigma, alpha = sp.alpha;
    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    for (iter = 1; iter <= sp.max_iter; ++iter) {
      // x~ = K^-1 rhs ; x <- alpha x~ + (1-alpha) x
      const double xt = kinv_matvec(sm, a, part);
      if (part == 0) {
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
          sm.rhs[padj(j)] = sigma * sm.x[j] - sm.qb[j] + s;
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
      if (tid < kN) sm.xD[padj(tid)] = sm.D[tid] * sm.x[tid];
      __syncthreads();
      double v[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = 0.0;
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
        const float4* prow = reinterpret_cast<const float4*>(&sm.P[r * kRowPad + part * kChunkPad]);
        const double2* xrow = reinterpret_cast<const double2*>(&sm.xD[part * kChunkPad]);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {  // pad lanes: P = 0, xD = 0
          const float4 pv = prow[j4];
          const double2 x0 = xrow[2 * j4], x1 = xrow[2 * j4 + 1];
          s0 = fma((double)pv.x, x0.x, s0);
          s1 = fma((double)pv.y, x0.y, s1);
          s0 = fma((double)pv.z, x1.x, s0);
          s1 = fma((double)pv.w, x1.y, s1);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0) {
          const double Px = sm.scal[0] * sm.D[r] * s;
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
        }
        __syncthreads();
        // rhs was built with the old rho vector: rebuild it
        if (crow) sm.w[ci] = sm.rv[ci] * sm.z[ci] - sm.y[ci];
        __syncwarp();
        if (crow && ce < 3) {
          const int j = 3 * ck + ce;
          const double* av = &sm.Av[9 * ck];
          const double* w = &sm.w[5 * ck];
          double s;
          if (ce == 0) s = av[0] * w[0] + av[2] * w[1];
          else if (ce == 1) s = av[4] * w[2] + av[6] * w[3];
          else s = av[1] * w[0] + av[3] * w[1] + av[5] * w[2] + av[7] * w[3] + av[8] * w[4];
          sm.rhs[padj(j)] = sigma * sm.x[j] - sm.qb[j] + s;
        }
        factor_inverse(sm, a, r, part, tid, sigma);
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    __syncthreads();
    if (x_all != nullptr && tid < kN) x_all[size_t(p) * kN + tid] = (float)(sm.D[tid] * sm.x[tid]);
    if (tid < 12) {
      const int leg = tid / 3, rr = tid % 3;
      const double f0 = sm.D[3 * leg] * sm.x[3 * leg];
      const double f1 = sm.D[3 * leg + 1] * sm.x[3 * leg + 1];
      const double f2 = sm.D[3 * leg + 2] * sm.x[3 * leg + 2];
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
