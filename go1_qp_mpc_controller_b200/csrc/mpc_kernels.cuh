// Device kernels of the batched convex-MPC GRF engine (sm_100a): shared constants and the QP build.
//
//   qp_build_kernel     K0+K1+K2  state record -> A_d, B_d, A_d powers, condensed Hessian
//                                 B_qp'QB_qp + R and gradient WITHOUT materialising B_qp
//                                 (S_j / T_j / U_jl regrouping, see the kernel), bounds
//                                 (ConvexMpc.cpp:110-245, A1RobotControl.cpp:452-518)
//   admm_solve_kernel   K3+K4+K5  (admm_kernel.cuh) OSQP-equivalent ADMM: Ruiz scaling,
//                                 K = P+sigma I+A'rho A, K^-1 by a blocked symmetric sweep held
//                                 in REGISTERS, ADMM loop, residual termination, rho adaptation,
//                                 unscale + R'f writer (+ fused torque map)
//                                 (OSQP 0.6.x as driven by A1RobotControl.cpp:522-561)
//
// Precision (measured on the CPU arithmetic model, see DESIGN.md "precision"):
// the GRF parity gate (1e-3 vs the fp64 oracle at eps 1e-5) needs the Hessian
// ACCUMULATED in fp64 and K, K^-1 and the ADMM iterates in fp64; fp32 is fine
// for the QP as read back through mpc_get_qp.  Between the two kernels P and q
// stay in f64 (HBM traffic is ~0.5 GB per 4096-state step, irrelevant next to the
// solve), which keeps the device iterate sequence identical to the oracle's:
// with f32 hand-over 0.3 % of states flipped a termination check.  B200 runs
// DFMA at half the FFMA rate (measured 17.1 T DFMA/s).
// P is padded to 128 columns per row (columns 120..127 exact zeros) so that one cp.async.bulk
// moves a whole problem into the solver's shared memory.
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
constexpr int kBuildThreads = 256;  // qp_build_kernel: two CTAs per SM (114 KB shared memory, 128 registers)
constexpr int kNP = 128;     // padded column count

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

// ---------------------------------------------------------------------------
// K0+K1+K2: QP build.  One CTA per problem, grid-stride over problems.
// ---------------------------------------------------------------------------
struct BuildSmem {
  double U[(kH * (kH - 1) / 2) * 156];  // A_d^(j-l) B_d(l), j > l: the sub-diagonal blocks of B_qp (56 KB)
  double CT[kH * 169];           // first C_m = (A^m)' Q A^m, then (C is dead) T_j = B_j' S_j
  double S[kH * 169];            // S_j = sum_{m <= H-1-j} C_m
  double g[kS];                  // gradient accumulators g_j
  double Apow[(kH + 1) * 169];   // A_d^0 .. A_d^H
  alignas(16) double Bd[kH * 156];  // B_mat_d_list (read as double2 by the Hessian items)
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

__global__ void __launch_bounds__(kBuildThreads, 2)
qp_build_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait, ModelIn model, int num,
                double* __restrict__ model_out, double* __restrict__ P_out,
                double* __restrict__ q_out, float* __restrict__ l_out, float* __restrict__ u_out,
                const __grid_constant__ BuildParams bp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BuildSmem& sm = *reinterpret_cast<BuildSmem*>(smem_raw);
  const int tid = threadIdx.x;

  for (int p = blockIdx.x; p < num; p += gridDim.x) {
    __syncthreads();  // smem reuse across problems
    if (states != nullptr) {
      // ---- K0: coalesced record load (48 floats) ----
      if (tid < 48) sm.st[tid] = reinterpret_cast<const float*>(states + p)[tid];
      __syncthreads();
      const float* st = sm.st;
      // ---- A_d = I + dt*A_c (ConvexMpc.cpp:110-130, :149-150) ----
      for (int vt = tid; vt < 480; vt += kBuildThreads) {  // the ranges below were laid out for 480 threads
        if (vt < 169) {
          const int rr = vt / 13, cc = vt % 13;
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
          sm.Apow[vt] = (rr == cc) ? 1.0 : 0.0;
          sm.Apow[169 + vt] = ((rr == cc) ? 1.0 : 0.0) + a * bp.dt;
        } else if (vt >= 192 && vt < 192 + 156) {
          for (int i = 0; i < kH; ++i) sm.Bd[i * 156 + vt - 192] = 0.0;
        } else if (vt >= 352 && vt < 352 + 13) {
          // mpc_states (A1RobotControl.cpp:452-456)
          const int k = vt - 352;
          sm.x0[k] = (k < 12) ? (double)st[k] : -9.8;
        } else if (vt >= 384 && vt < 384 + kH) {
          // mpc_states_d step i (A1RobotControl.cpp:470-488)
          const int i = vt - 384;
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
        } else if (vt >= 416 && vt < 416 + 4 * kH) {
          const int i = (vt - 416) >> 2, leg = (vt - 416) & 3;
          int c = st[kOffContacts + leg] != 0.0f;
          if (bp.gait_aware && i > 0) {
            // planned contact of step i from the gait counter (A1RobotControl.cpp:156-164)
            const float* g = reinterpret_cast<const float*>(gait + p);
            const double cnt = fmod((double)g[leg] + (double)i * (double)g[10] * (double)g[4 + leg], (double)g[8]);
            c = cnt <= (double)g[9];
          }
          sm.contacts[4 * i + leg] = c;
        }
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
        for (int idx = tid; idx < (kH - 1) * 156; idx += kBuildThreads) sm.Bd[156 + idx] = sm.Bd[idx % 156];
    } else {
      // ---- ConvexMpc surface: caller-written A_mat_d / B_mat_d_list ----
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = model.A_d[size_t(p) * 169 + tid];
      }
      for (int idx = tid; idx < kH * 156; idx += kBuildThreads)
        sm.Bd[idx] = model.B_d_list[size_t(p) * kH * 156 + idx];
      if (tid < 13) sm.x0[tid] = model.x0[size_t(p) * 13 + tid];
      if (tid < kS) sm.xref[tid] = model.x_ref[size_t(p) * kS + tid];
      if (tid < 4 * kH) sm.contacts[tid] = model.contacts[size_t(p) * 4 + (tid & 3)] != 0;
    }
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
    if (model_out != nullptr) {
      // A_d and the B_d list for the structured (Riccati) solver
      double* mo = model_out + size_t(p) * (169 + kH * 156);
      for (int idx = tid; idx < 169 + kH * 156; idx += kBuildThreads)
        mo[idx] = (idx < 169) ? sm.Apow[169 + idx] : sm.Bd[idx - 169];
    }
    // ---- K2: the condensed Hessian WITHOUT materialising B_qp (ConvexMpc.cpp:185-217).
    // B_qp block (i, j) = A_d^(i-j) B_d(j) for i >= j, so with S_j = sum_{m <= H-1-j} (A^m)' Q A^m
    //   Hessian block (j, l), j >= l:  B_j' S_j A^(j-l) B_l = T_j U_jl   (+ R on the diagonal)
    //   gradient block j:              B_j' g_j,  g_j = sum_{i >= j} (A^(i-j))' Q (A^(i+1) x0 - x_ref,i)
    // 0.27 M multiply-adds per problem instead of 1.9 M, no 133 KB B_qp in shared memory; the
    // sums run over the same products as B_qp' Q B_qp, regrouped (difference ~1e-16 relative).
    // C_m = (A^m)' Q A^m, m < H ; tmp_i = Q (A^(i+1) x0 - x_ref,i)
    for (int idx = tid; idx < kH * 169; idx += kBuildThreads) {
      const int m = idx / 169, e = idx - 169 * m, ia = e / 13, ib = e - 13 * ia;
      const double* Am = &sm.Apow[m * 169];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) s = fma(Am[k * 13 + ia] * bp.Qd[k], Am[k * 13 + ib], s);
      sm.CT[idx] = s;
    }
    if (tid < kS) {
      const int i = tid / 13, rr = tid % 13;
      const double* Ai = &sm.Apow[(i + 1) * 169 + rr * 13];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) s += Ai[k] * sm.x0[k];
      sm.tmp[tid] = bp.Qd[rr] * (s - sm.xref[tid]);
    }
    __syncthreads();
    // S_j = C_0 + .. + C_(H-1-j) (running sum, one thread per entry) ; g_j (one thread per (j, a))
    if (tid < 169) {
      double acc = 0.0;
      for (int m = 0; m < kH; ++m) {
        acc += sm.CT[m * 169 + tid];
        sm.S[(kH - 1 - m) * 169 + tid] = acc;
      }
    }
    if (tid >= kBuildThreads - kS) {
      const int t = tid - (kBuildThreads - kS), j = t / 13, ia = t % 13;
      double s = 0.0;
      for (int i = j; i < kH; ++i) {
        const double* Am = &sm.Apow[(i - j) * 169];
#pragma unroll
        for (int k = 0; k < 13; ++k) s = fma(Am[k * 13 + ia], sm.tmp[13 * i + k], s);
      }
      sm.g[t] = s;
    }
    __syncthreads();
    // T_j = B_j' S_j (12 x 13) ; U_jl = A^(j-l) B_l (13 x 12), j > l ; gradient = B_j' g_j
    for (int idx = tid; idx < kH * 156; idx += kBuildThreads) {
      const int j = idx / 156, e = idx - 156 * j, ia = e / 13, k = e - 13 * ia;
      const double* Bj = &sm.Bd[j * 156];
      const double* Sj = &sm.S[j * 169];
      double s = 0.0;
#pragma unroll
      for (int pp = 0; pp < 13; ++pp) s = fma(Bj[pp * 12 + ia], Sj[pp * 13 + k], s);
      sm.CT[idx] = s;
    }
    for (int idx = tid; idx < (kH * (kH - 1) / 2) * 156; idx += kBuildThreads) {
      const int blk = idx / 156, e = idx - 156 * blk, rr = e / 12, cc = e - 12 * rr;
      // blk -> (j, l) with l < j, row-major over the strict lower triangle
      int j = 1, rem = blk;
      while (rem >= j) { rem -= j; ++j; }
      const int l = rem;
      const double* Ap = &sm.Apow[(j - l) * 169];
      const double* Bl = &sm.Bd[l * 156];
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) s = fma(Ap[rr * 13 + k], Bl[k * 12 + cc], s);
      sm.U[idx] = s;
    }
    if (tid < kN) {
      const int j = tid / 12, ia = tid - 12 * j;
      const double* Bj = &sm.Bd[j * 156];
      double s = 0.0;
#pragma unroll
      for (int pp = 0; pp < 13; ++pp) s = fma(Bj[pp * 12 + ia], sm.g[13 * j + pp], s);
      q_out[size_t(p) * kN + tid] = s;
    }
    __syncthreads();
    // Hessian: one work item per 6 x 6 quarter of a block (j, l), j >= l (220 items).  The item
    // holds six rows of T_j and six columns of U_jl in registers (78 LDS.128 for 468 FMAs) and
    // writes its quarter AND the mirrored one, so the matrix is exactly symmetric.
    for (int item = tid; item < 4 * (kH * (kH + 1) / 2); item += kBuildThreads) {
      const int blk = item >> 2, sa = (item >> 1) & 1, sb = item & 1;
      int j = 0, rem = blk;
      while (rem > j) { rem -= (j + 1); ++j; }
      const int l = rem;                                  // block (j, l), l <= j
      const double* Tj = &sm.CT[j * 156 + 6 * sa * 13];     // rows 6 sa .. of T_j, contiguous
      const double* Ujl = (j == l) ? &sm.Bd[j * 156] : &sm.U[(j * (j - 1) / 2 + l) * 156];
      double t[6][13];
      {
        const double2* tp = reinterpret_cast<const double2*>(Tj);
        double flat[78];
#pragma unroll
        for (int h = 0; h < 39; ++h) { const double2 v = tp[h]; flat[2 * h] = v.x; flat[2 * h + 1] = v.y; }
#pragma unroll
        for (int ra = 0; ra < 6; ++ra)
#pragma unroll
          for (int k = 0; k < 13; ++k) t[ra][k] = flat[13 * ra + k];
      }
      double acc[6][6];
#pragma unroll
      for (int ra = 0; ra < 6; ++ra)
#pragma unroll
        for (int cb = 0; cb < 6; ++cb) acc[ra][cb] = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) {
        const double2* up = reinterpret_cast<const double2*>(Ujl + k * 12 + 6 * sb);
        const double2 u01 = up[0], u23 = up[1], u45 = up[2];
        const double u[6] = {u01.x, u01.y, u23.x, u23.y, u45.x, u45.y};
#pragma unroll
        for (int ra = 0; ra < 6; ++ra)
#pragma unroll
          for (int cb = 0; cb < 6; ++cb) acc[ra][cb] = fma(t[ra][k], u[cb], acc[ra][cb]);
      }
      if (j == l) {
        // diagonal block: symmetrise by its lower triangle, add R (ConvexMpc.cpp:211)
        if (sa == sb) {
#pragma unroll
          for (int ra = 0; ra < 6; ++ra) {
#pragma unroll
            for (int cb = ra + 1; cb < 6; ++cb) acc[ra][cb] = acc[cb][ra];
            acc[ra][ra] += bp.Rd[6 * sa + ra];
          }
        } else if (sa < sb) {
          continue;  // quarter (0, 1) of a diagonal block is written by the item of quarter (1, 0)
        }
      }
      // P is handed to the solver in f64, rows padded to 128 (pad columns are exact zeros),
      // so one cp.async.bulk moves a whole problem into shared memory
      double* Pb = P_out + size_t(p) * kN * kNP;
#pragma unroll
      for (int ra = 0; ra < 6; ++ra) {
        double2* dst = reinterpret_cast<double2*>(Pb + size_t(12 * j + 6 * sa + ra) * kNP + 12 * l + 6 * sb);
        dst[0] = make_double2(acc[ra][0], acc[ra][1]);
        dst[1] = make_double2(acc[ra][2], acc[ra][3]);
        dst[2] = make_double2(acc[ra][4], acc[ra][5]);
      }
      if (j != l || sa != sb) {
        // the mirrored quarter: rows of block l, columns of block j
#pragma unroll
        for (int cb = 0; cb < 6; ++cb) {
          double2* dst = reinterpret_cast<double2*>(Pb + size_t(12 * l + 6 * sb + cb) * kNP + 12 * j + 6 * sa);
          dst[0] = make_double2(acc[0][cb], acc[1][cb]);
          dst[1] = make_double2(acc[2][cb], acc[3][cb]);
          dst[2] = make_double2(acc[4][cb], acc[5][cb]);
        }
      }
    }
    // zero padding of the rows (columns 120..127)
    if (tid < kN) {
      double2* dst = reinterpret_cast<double2*>(P_out + size_t(p) * kN * kNP + size_t(tid) * kNP + kN);
#pragma unroll
      for (int h = 0; h < (kNP - kN) / 2; ++h) dst[h] = make_double2(0.0, 0.0);
    }
    // ---- bounds, contacts replicated over the horizon (ConvexMpc.cpp:223-245) ----
    for (int i = tid; i < kM; i += kBuildThreads) {
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
