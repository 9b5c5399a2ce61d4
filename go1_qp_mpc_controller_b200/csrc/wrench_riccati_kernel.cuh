// K0..K5 fused, long horizon (H = 30): the whole compute_grf MPC branch (A1RobotControl.cpp:446-561) for
// one robot state per CTA -- state record in, body-frame GRF (+ joint torques) out -- with NO Hessian and no
// model in memory.  One CTA of 256 threads per problem, TWO CTAs per SM, problems pulled from an atomic
// counter.  Same OSQP 0.6.x iteration as every other solver of this library (Ruiz equilibration, per-row
// rho, alpha relaxation, checks every 25, rho adaptation every 50); what changes is the linear algebra.
//
// Structure (wrench_kernel.cuh has the derivation; scripts/proto_wrench_riccati.py checks this file's
// formulas against the dense ones -- identical iteration counts, GRF within 1e-11):
//   K = c D P D + sigma I + A_' rho A_ = G_' C G_ + Delta,  G_ block diagonal 6 x 12 per step, Delta block
//   diagonal 3 x 3 per leg-step, C = c S the 6H x 6H wrench-space Hessian, and by Woodbury
//       K^-1 r = a - M~' (N u - dlt),  a = Delta^-1 r,  u = M~ r,  M~ = N^-1 G_ Delta^-1,  N = G_ Delta^-1 G_',
//       dlt = (C + N^-1)^-1 u.
//   At H = 10 that core is a dense 60 x 60 matrix in registers.  Here it is what it also is: the Hessian of an
//   LQR problem with SIX inputs per step (the velocity increment dlt_k of the step) on TWELVE states
//   X = (euler, position | angular, linear velocity),  X_k+1 = A X_k + [0; I] dlt_k,  A = [[I, dt Rt], [0, I]]:
//     factor (per rho):  Pi = cQ;  k = H-1 .. 0:  Z_k = L (I + L' Pi_vv L)^-1 L'  (N_k = L L'),  U = (Pi A)_v,
//                        F_k = -Z_k U,  Pi <- cQ + A' Pi A + U' F_k
//     solve (per iteration):  p_k = (A + [0; I] F_k)' p_k+1 - F_k' u_k  (backward),  e_k = p_k+1,v - u_k,
//                             dlt_k = F_k X_k - Z_k e_k,  X_k+1 = A X_k + [0; I] dlt_k  (forward)
//   -- 12 x 6 per step instead of the 13 x 12 of riccati_kernel.cuh, 60 KB of per-step matrices instead of
//   150 KB, so that two problems share an SM, and no dense Hessian for the equilibration either (column
//   norms in closed form, S_kl = alpha_kl D1 + beta_kl D2).
//
// Threads: warp w, team t = lane / 8, t8 = lane % 8; the team owns horizon step k = 4 w + t (teams 30, 31 idle).
//   leg role   t8 < 4: leg t8 of the step -- its three variables, five constraint rows (z, y in registers,
//              normalised as in wrench_kernel.cuh), the 3 x 3 block of Delta^-1.
//   axis role  t8 < 6: component t8 of the step's six-vectors (u, e, dlt) and the state pair (pos_c, vel_c).
// All per-step phases exchange through shared memory inside the team (__syncwarp only).  The two recursions
// run as three sweeps: every warp over its own four steps from a zero boundary, warp 0 over the eight group
// boundaries through Phi_j = Acl_(4j+3) ... Acl_4j, every warp again from its true boundary: 4 + 6 + 4
// dependent steps of one 12-term dot product each instead of 30, and four block barriers per iteration.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "wrench_kernel.cuh"

namespace mpcb200 {

constexpr int kWrcThreads = 256;
constexpr int kWrcWarps = kWrcThreads / 32;
constexpr int kWrcCtasPerSm = 2;
constexpr int kMS = 14;  // row stride of the 6 x 12 and 12 x 12 matrices (16-byte loads of six rows hit six bank groups)

template <int H>
struct WrcSmem {
  static constexpr int n = 12 * H, m = 20 * H, nG = (H + 3) / 4;
  alignas(16) double Mt[H][6 * kMS];   // M~ = N^-1 G_ Delta^-1 (rows; first G_ Delta^-1 during a factorisation)
  alignas(16) double Fk[H][6 * kMS];   // F_k (rows; first G_ during a factorisation)
  alignas(16) double Lk[H][36];        // Cholesky factor of N_k, row-major, zeros above the diagonal
  alignas(16) double Nk[H][36];
  alignas(16) double Zk[H][36];
  double Li[H][6];                     // 1 / L_cc (0 for a zero pivot)
  alignas(16) double Phi[nG - 2][12 * kMS];  // Phi_1 .. Phi_(nG-2)
  alignas(16) double rhs[n];
  alignas(16) double uv[6 * H];
  alignas(16) double ev[6 * H];        // e, then h = N u - dlt
  alignas(16) double pv[(H + 1) * 12];
  alignas(16) double Xv[(H + 1) * 12];
  alignas(16) double Pb[(nG + 1) * 12];  // group boundaries of the backward recursion (index = group)
  alignas(16) double Xb[(nG + 1) * 12];  // ... of the forward recursion
  alignas(16) double zero12[12];
  alignas(16) double Dp[n];
  alignas(16) double cca[m];           // E_row D_own of every row (checks, warm slot)
  alignas(16) double gam[6 * H];
  // Ruiz: column-norm halves [2][n] | build: Q e [H][14] | factorisation: Riccati scratch | check: D x, G D x, S G D x
  alignas(16) double scr[2 * n];
  alignas(16) double B6t[3][12];       // top rows of B6c (step 0)
  double dT[9];                        // foot_drift: top rows of step k are B6t - k dT (same for every leg)
  double red[kWrcWarps * 16];
  double scal[16];                     // 0:c 1:1/c 2:rho 4:pri_res
  float be[H * H];                     // sum_{i >= max(k,l)} (i - k)(i - l)  (exact in fp32)
  float st[48];
  int contacts[4 * H];
  int flags[8];                        // 0:done 1:status 2:refactor 3:problem index
};

// max_i |(G' S G)_ij| D_i over the rows i of two legs (6 hp .. 6 hp + 5 of every step), for column j of
// horizon step kj, component comp:  (G' S G)_ij = top_i . v + [comp_i == comp] vbm,  v = alpha u1 + beta u2
template <int H, bool kDrift>
__device__ __forceinline__ double wrc_colnorm_t(const WrcSmem<H>& sm, int hp, int kj, int comp, const double (&u1)[3],
                                                const double (&u2)[3], double vb1, double vb2, double dt2, double dt4) {
  double mx = 0.0;
  double tp[3][6];
#pragma unroll
  for (int c3 = 0; c3 < 3; ++c3) {
    const double2* t = reinterpret_cast<const double2*>(&sm.B6t[c3][6 * hp]);
    const double2 q0 = t[0], q1 = t[1], q2 = t[2];
    tp[c3][0] = q0.x; tp[c3][1] = q0.y; tp[c3][2] = q1.x; tp[c3][3] = q1.y; tp[c3][4] = q2.x; tp[c3][5] = q2.y;
  }
  double dT[9];
  if (kDrift) {
#pragma unroll
    for (int i = 0; i < 9; ++i) dT[i] = sm.dT[i];
  }
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const int mxk = k > kj ? k : kj;
    const double a = (double)(H - mxk) * dt2, b = dt4 * (double)sm.be[H * k + kj];
    const double v0 = fma(b, u2[0], a * u1[0]), v1 = fma(b, u2[1], a * u1[1]), v2 = fma(b, u2[2], a * u1[2]);
    const double vbm = fma(b, vb2, a * vb1);
    const double2* dk = reinterpret_cast<const double2*>(&sm.Dp[12 * k + 6 * hp]);
    const double2 d01 = dk[0], d23 = dk[1], d45 = dk[2];
    const double dd[6] = {d01.x, d01.y, d23.x, d23.y, d45.x, d45.y};
    const double kd = (double)k;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double e = ((i % 3) == comp) ? vbm : 0.0;   // (6 hp + i) % 3 == i % 3
      if (kDrift) {
        e = fma(fma(-kd, dT[i % 3], tp[0][i]), v0, e);
        e = fma(fma(-kd, dT[3 + i % 3], tp[1][i]), v1, e);
        e = fma(fma(-kd, dT[6 + i % 3], tp[2][i]), v2, e);
      } else {
        e = fma(tp[0][i], v0, e);
        e = fma(tp[1][i], v1, e);
        e = fma(tp[2][i], v2, e);
      }
      mx = max_bits(mx, fabs(e) * dd[i]);
    }
  }
  return mx;
}

// block-wide sum (slot 0) and max (slot 1) of one value pair per thread; result to every thread
__device__ __forceinline__ void wrc_block_sum_max(double& s, double& q, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q = fmax(q, __shfl_xor_sync(0xffffffffu, q, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[2 * (threadIdx.x >> 5)] = s;
    red[2 * (threadIdx.x >> 5) + 1] = q;
  }
  __syncthreads();
  s = 0.0;
  q = 0.0;
#pragma unroll
  for (int w = 0; w < kWrcWarps; ++w) {
    s += red[2 * w];
    q = fmax(q, red[2 * w + 1]);
  }
}

template <int H>
__global__ void __launch_bounds__(kWrcThreads, kWrcCtasPerSm)
wrench_riccati_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait,
                      MpcResult* __restrict__ results, float* __restrict__ x_all, int num, int* __restrict__ counter,
                      double* __restrict__ warm, int warm_stride, const MpcTorqueIn* __restrict__ tin,
                      MpcTorqueOut* __restrict__ tout, const __grid_constant__ BuildParams bp,
                      const __grid_constant__ SolveParams sp) {
  using Smem = WrcSmem<H>;
  constexpr int n = Smem::n, m = Smem::m, nG = Smem::nG;
  constexpr int kWX = 0, kWQ = n, kWZ = 2 * n, kWY = 2 * n + m, kWRho = 2 * n + 2 * m, kWLive = kWRho + 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = lane >> 3, t8 = lane & 7;
  const int kraw = 4 * warp + team;
  const bool on = kraw < H;
  const int k = on ? kraw : H - 1;                      // idle teams shadow the last step (never write)
  const bool isleg = on && t8 < 4, isax = on && t8 < 6;
  const int lg = t8 & 3;                                // leg (leg role)
  const int c = t8 < 6 ? t8 : t8 - 6;                   // component (axis role)
  const int j0 = 12 * k + 3 * lg, r0 = 5 * (4 * k + lg);  // first variable / first row of the leg-step
  const int klast = (4 * warp + 3 < H - 1) ? 4 * warp + 3 : H - 1;  // last step of this warp's group
  const bool kWarm = warm != nullptr;
  const double mu = sp.mu, sigma = sp.sigma, alpha = sp.alpha;
  const double dt = bp.dt, inv_m = 1.0 / bp.mass;
  const double dt2 = dt * dt, dt4 = dt2 * dt2;
  const double r2x = bp.Rd[3 * lg], r2y = bp.Rd[3 * lg + 1], r2z = bp.Rd[3 * lg + 2];

  // ---- once per CTA ----
  for (int i = tid; i < H * H; i += kWrcThreads) {
    const int kk = i / H, ll = i - H * kk;
    const int mxk = kk > ll ? kk : ll;
    int s = 0;
    for (int q = mxk; q < H; ++q) s += (q - kk) * (q - ll);
    sm.be[i] = (float)s;
  }
  if (tid < 12) {
    sm.zero12[tid] = 0.0;
    sm.Xb[tid] = 0.0;  // X_0 = 0: the true boundary of the first group
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;

    // ---- K0: record load ----
    if (tid < 48) sm.st[tid] = reinterpret_cast<const float*>(states + p)[tid];
    double* const ws = kWarm ? warm + size_t(p) * warm_stride : nullptr;
    const bool live = kWarm && ws[kWLive] != 0.0;
    const double rho0 = live ? ws[kWRho] : sp.rho;
    if (tid == 0) {
      sm.scal[0] = 1.0;
      sm.scal[2] = rho0;
      sm.flags[0] = 0;
      sm.flags[1] = MPC_STATUS_UNSOLVED;
    }
    __syncthreads();
    const float* st = sm.st;

    // ---- K1: model in closed form (ConvexMpc.cpp:110-156, A1RobotControl.cpp:452-514) ----
    double syaw, cyaw;
    sincos((double)st[kOffEuler + 2], &syaw, &cyaw);
    const double Q0 = bp.Qd[0], Q1 = bp.Qd[1], Q2 = bp.Qd[2];
    const double th00 = cyaw * cyaw * Q0 + syaw * syaw * Q1, th01 = cyaw * syaw * Q0 - syaw * cyaw * Q1,
                 th11 = syaw * syaw * Q0 + cyaw * cyaw * Q1, th22 = Q2;
    double (*Qe)[14] = reinterpret_cast<double (*)[14]>(sm.scr);
    if (tid < H) {
      // Q (A_d^(i+1) x0 - x_ref,i): A_d^m x0 = x0 + m dt A_c x0 + c2 g e_5 (A_c^2 x0 = g e_5, A_c^3 = 0)
      const int i = tid;
      const double mm = (double)(i + 1);
      const double gr = -9.8;
      const double c2 = 0.5 * mm * (mm - 1.0) * dt * dt;
      const double wx = st[kOffAngVel], wy = st[kOffAngVel + 1], wz = st[kOffAngVel + 2];
      const double vx = st[kOffLinVel], vy = st[kOffLinVel + 1], vz = st[kOffLinVel + 2];
      const double R0 = st[kOffRot + 0], R1 = st[kOffRot + 1], R2 = st[kOffRot + 2];
      const double R3 = st[kOffRot + 3], R4 = st[kOffRot + 4], R5 = st[kOffRot + 5];
      const double dx = st[kOffLinVelD], dy = st[kOffLinVelD + 1], dz = st[kOffLinVelD + 2];
      const double vwx = R0 * dx + R1 * dy + R2 * dz, vwy = R3 * dx + R4 * dy + R5 * dz;  // :470
      double xi[13], xr[13];
      xi[0] = (double)st[kOffEuler] + mm * dt * (cyaw * wx + syaw * wy);
      xi[1] = (double)st[kOffEuler + 1] + mm * dt * (-syaw * wx + cyaw * wy);
      xi[2] = (double)st[kOffEuler + 2] + mm * dt * wz;
      xi[3] = (double)st[kOffPos] + mm * dt * vx;
      xi[4] = (double)st[kOffPos + 1] + mm * dt * vy;
      xi[5] = (double)st[kOffPos + 2] + mm * dt * vz + c2 * gr;
      xi[6] = wx; xi[7] = wy; xi[8] = wz;
      xi[9] = vx; xi[10] = vy; xi[11] = vz + mm * dt * gr;
      xi[12] = gr;
      xr[0] = (double)st[kOffEulerD];                                             // :472-488
      xr[1] = (double)st[kOffEulerD + 1];
      xr[2] = (double)st[kOffEuler + 2] + (double)st[kOffAngVelD + 2] * dt * mm;
      xr[3] = (double)st[kOffPos] + vwx * dt * mm;
      xr[4] = (double)st[kOffPos + 1] + vwy * dt * mm;
      xr[5] = (double)st[kOffPosDz];
      xr[6] = (double)st[kOffAngVelD]; xr[7] = (double)st[kOffAngVelD + 1]; xr[8] = (double)st[kOffAngVelD + 2];
      xr[9] = vwx; xr[10] = vwy; xr[11] = 0.0; xr[12] = gr;
#pragma unroll
      for (int q = 0; q < 13; ++q) Qe[i][q] = bp.Qd[q] * (xi[q] - xr[q]);
    } else if (tid >= 32 && tid < 37) {
      // top rows of B6c, I_w^-1 [r_leg]x, per leg (ConvexMpc.cpp:132-143); thread 36: their drift per step
      const int t = tid - 32;
      double R[9], T[9], Iw[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = (double)st[kOffRot + i];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          double s = 0.0;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) s += R[3 * i + kk] * bp.inertia[3 * kk + q];
          T[3 * i + q] = s;
        }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          double s = 0.0;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) s += T[3 * i + kk] * R[3 * q + kk];
          Iw[3 * i + q] = s;
        }
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
      double fx, fy, fz;
      if (t < 4) {
        fx = st[kOffFoot + 3 * t]; fy = st[kOffFoot + 3 * t + 1]; fz = st[kOffFoot + 3 * t + 2];
      } else {
        // foot positions of step k are r - k dt R v_d (A1RobotControl.cpp:504-507, commented out there)
        const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
        fx = dt * (R[0] * vx + R[1] * vy + R[2] * vz);
        fy = dt * (R[3] * vx + R[4] * vy + R[5] * vz);
        fz = dt * (R[6] * vx + R[7] * vy + R[8] * vz);
      }
      const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};  // Utils::skew
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          double s = 0.0;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) s += Inv[3 * i + kk] * sk[3 * kk + q];
          if (t < 4) sm.B6t[i][3 * t + q] = s;
          else sm.dT[3 * i + q] = bp.foot_drift ? s : 0.0;
        }
    } else if (tid >= 64 && tid < 64 + 4 * H) {
      const int t = tid - 64, stp = t >> 2, l4 = t & 3;
      int cf = st[kOffContacts + l4] != 0.0f;
      if (bp.gait_aware && stp > 0) {
        // planned contact of step i from the gait counter (A1RobotControl.cpp:156-164)
        const float* gi = reinterpret_cast<const float*>(gait + p);
        const double cnt = fmod((double)gi[l4] + (double)stp * (double)gi[10] * (double)gi[4 + l4], (double)gi[8]);
        cf = cnt <= (double)gi[9];
      }
      sm.contacts[t] = cf;
    }
    for (int i = tid; i < n; i += kWrcThreads) sm.Dp[i] = 1.0;
    __syncthreads();
    // gam_k = sum_{i >= k} F_(i-k)' Q e_i  (one thread per wrench component)
    if (tid < 6 * H) {
      const int kk = tid / 6, cc = tid - 6 * kk;
      double s = 0.0;
      for (int i = kk; i < H; ++i) {
        const double kp = (double)(i - kk) * dt * dt;
        const double* e = Qe[i];
        double lin, rot;
        if (cc == 0) { lin = e[6]; rot = cyaw * e[0] - syaw * e[1]; }
        else if (cc == 1) { lin = e[7]; rot = syaw * e[0] + cyaw * e[1]; }
        else if (cc == 2) { lin = e[8]; rot = e[2]; }
        else { lin = e[6 + cc]; rot = e[cc]; }
        s += dt * lin + kp * rot;
      }
      sm.gam[tid] = s;
    }
    __syncthreads();

    // ---- per-leg constants of the build (leg role) ----
    const bool drift = bp.foot_drift != 0;
    double top[3][3];  // top[i][q]: row i of B6c, column 3 lg + q, at this step
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int q = 0; q < 3; ++q) top[i][q] = fma(-(double)k, sm.dT[3 * i + q], sm.B6t[i][3 * lg + q]);
    double q0v[3], pjj[3];
    {
      const double a = (double)(H - k) * dt2, b = dt4 * (double)sm.be[H * k + k];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        // gradient (ConvexMpc.cpp:215-217): q_j = B6c[:, j] . gam_k
        q0v[q] = top[0][q] * sm.gam[6 * k] + top[1][q] * sm.gam[6 * k + 1] + top[2][q] * sm.gam[6 * k + 2] +
                 inv_m * sm.gam[6 * k + 3 + q];
        // diagonal entry of P (the only one R2 touches): P_jj = B6c_j' S_kk B6c_j + 2 r_j
        const double u10 = bp.Qd[6] * top[0][q], u11 = bp.Qd[7] * top[1][q], u12 = bp.Qd[8] * top[2][q];
        const double u20 = th00 * top[0][q] + th01 * top[1][q], u21 = th01 * top[0][q] + th11 * top[1][q],
                     u22 = th22 * top[2][q];
        const double vb1 = bp.Qd[9 + q] * inv_m * inv_m, vb2 = bp.Qd[3 + q] * inv_m * inv_m;
        pjj[q] = top[0][q] * (a * u10 + b * u20) + top[1][q] * (a * u11 + b * u21) + top[2][q] * (a * u12 + b * u22) +
                 (a * vb1 + b * vb2) + bp.Rd[3 * lg + q];
      }
    }
    double qsc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) qsc[q] = (live && isleg) ? ws[kWQ + j0 + q] : q0v[q];
    // bounds of row 4 (ConvexMpc.cpp:223-245); fp32 like the dense path's hand-over.  Rows 0, 2 are [0, inf),
    // rows 1, 3 (-inf, 0]: only their finite side can bind.
    double lo4, hi4;
    {
      const float cflag = sm.contacts[4 * k + lg] ? 1.0f : 0.0f;
      lo4 = (double)((float)bp.fz_min * cflag);
      hi4 = (double)((float)bp.fz_max * cflag);
    }

    // column-norm items of the equilibration: (column j, leg pair hp), three per thread
    auto colnorm_pass = [&]() {
      double* Px = sm.scr;
#pragma unroll 1
      for (int it = tid; it < 2 * n; it += kWrcThreads) {
        const int j = it >> 1, hp = it & 1;
        const int kj = j / 12, jj = j - 12 * kj, comp = jj % 3;
        double tj[3], u1[3], u2[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) tj[i] = fma(-(double)kj, sm.dT[3 * i + comp], sm.B6t[i][jj]);
        u1[0] = bp.Qd[6] * tj[0]; u1[1] = bp.Qd[7] * tj[1]; u1[2] = bp.Qd[8] * tj[2];
        u2[0] = th00 * tj[0] + th01 * tj[1]; u2[1] = th01 * tj[0] + th11 * tj[1]; u2[2] = th22 * tj[2];
        const double Qv = (comp == 0) ? bp.Qd[9] : (comp == 1) ? bp.Qd[10] : bp.Qd[11];
        const double Qp = (comp == 0) ? bp.Qd[3] : (comp == 1) ? bp.Qd[4] : bp.Qd[5];
        const double vb1 = Qv * inv_m * inv_m, vb2 = Qp * inv_m * inv_m;
        Px[n * hp + j] = drift ? wrc_colnorm_t<H, true>(sm, hp, kj, comp, u1, u2, vb1, vb2, dt2, dt4)
                               : wrc_colnorm_t<H, false>(sm, hp, kj, comp, u1, u2, vb1, vb2, dt2, dt4);
      }
    };

    // ---- K3a: modified Ruiz equilibration (osqp scaling.c scale_data), nothing materialised ----
    double D[3] = {1.0, 1.0, 1.0}, E[5] = {1.0, 1.0, 1.0, 1.0, 1.0}, c_run = 1.0;
    if (sp.scaling > 0) {
      double nP[3];
      colnorm_pass();
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 3; ++q) nP[q] = max_bits(max_bits(sm.scr[j0 + q], sm.scr[n + j0 + q]), pjj[q]);
      __syncthreads();  // scr and Dp are rewritten inside the loop
      for (int it = 0; it < sp.scaling; ++it) {
        // column norms of [P; A] and row norms of A from the current D, E
        const double nAx = fmax(E[0], E[1]) * D[0], nAy = fmax(E[2], E[3]) * D[1];
        const double nAz = fmax(mu * fmax(fmax(E[0], E[1]), fmax(E[2], E[3])), E[4]) * D[2];
        const double Dnx = D[0] * rsqrt(limit_scaling(fmax(nP[0], nAx)));
        const double Dny = D[1] * rsqrt(limit_scaling(fmax(nP[1], nAy)));
        const double Dnz = D[2] * rsqrt(limit_scaling(fmax(nP[2], nAz)));
        const double nrx = fmax(D[0], mu * D[2]), nry = fmax(D[1], mu * D[2]), nrz = D[2];
        E[0] = E[0] * rsqrt(limit_scaling(E[0] * nrx));
        E[1] = E[1] * rsqrt(limit_scaling(E[1] * nrx));
        E[2] = E[2] * rsqrt(limit_scaling(E[2] * nry));
        E[3] = E[3] * rsqrt(limit_scaling(E[3] * nry));
        E[4] = E[4] * rsqrt(limit_scaling(E[4] * nrz));
        D[0] = Dnx; D[1] = Dny; D[2] = Dnz;
        if (isleg) { sm.Dp[j0] = D[0]; sm.Dp[j0 + 1] = D[1]; sm.Dp[j0 + 2] = D[2]; }
        __syncthreads();
        colnorm_pass();
        __syncthreads();
        // cost normalisation with the new D and the old c
        const double c_old = c_run;
        double nP2[3], part_sum = 0.0, part_q = 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          nP2[q] = c_old * D[q] * max_bits(max_bits(sm.scr[j0 + q], sm.scr[n + j0 + q]), pjj[q] * D[q]);
          if (isleg) {
            part_sum += nP2[q];
            part_q = fmax(part_q, fabs(c_old * D[q] * qsc[q]));
          }
        }
        wrc_block_sum_max(part_sum, part_q, sm.red + (it & 1) * (2 * kWrcWarps));
        const double mean = part_sum / (double)n;
        const double ct = 1.0 / limit_scaling(fmax(mean, limit_scaling(part_q)));
        c_run = c_old * ct;
#pragma unroll
        for (int q = 0; q < 3; ++q) nP[q] = nP2[q] * ct;
      }
    }
    const double cs = c_run;
    if (tid == 0) { sm.scal[0] = cs; sm.scal[1] = 1.0 / cs; }
    double qb[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) qb[q] = cs * D[q] * q0v[q];
    // constraint types (auxil.c set_rho_vec): -1 loose, 1 equality, 0 inequality
    auto ctype_of = [](double lo, double hi) {
      return (lo < -MPC_INFTY * 1e-4 && hi > MPC_INFTY * 1e-4) ? -1 : ((hi - lo < 1e-4) ? 1 : 0);
    };
    auto rho_of = [](int ct, double rho) { return (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho; };
    int ct[5];
    ct[0] = ctype_of(0.0, (double)(float)MPC_INFTY * E[0]);
    ct[1] = ctype_of(-(double)(float)MPC_INFTY * E[1], 0.0);
    ct[2] = ctype_of(0.0, (double)(float)MPC_INFTY * E[2]);
    ct[3] = ctype_of(-(double)(float)MPC_INFTY * E[3], 0.0);
    ct[4] = ctype_of(lo4 * E[4], hi4 * E[4]);
    // Rows in normalised form (wrench_kernel.cuh): row i = cca_i (x_own +- t x_fz), the kernel iterates on
    // zh = z / cca and uh = y / (rho cca);  kap = rho cca^2.
    double cca[5], kap[5], rv[5];
    cca[0] = E[0] * D[0]; cca[1] = E[1] * D[0]; cca[2] = E[2] * D[1]; cca[3] = E[3] * D[1]; cca[4] = E[4] * D[2];
    const double tzx = mu * D[2] / D[0], tzy = mu * D[2] / D[1];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      rv[i] = rho_of(ct[i], rho0);
      kap[i] = rv[i] * cca[i] * cca[i];
      if (isleg) sm.cca[r0 + i] = cca[i];
    }
    lo4 = lo4 * E[4] / cca[4];
    hi4 = hi4 * E[4] / cca[4];

    // iterates
    double x[3] = {0.0, 0.0, 0.0}, z[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, u[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (live && isleg) {
#pragma unroll
      for (int q = 0; q < 3; ++q) x[q] = ws[kWX + j0 + q];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        z[i] = ws[kWZ + r0 + i] / cca[i];
        u[i] = ws[kWY + r0 + i] / (rv[i] * cca[i]);
      }
    }
    double di[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // Delta_leg^-1: 00 01 02 11 12 22
    double rh[3];                                    // rhs of the leg's variables

    // rhs = sigma x - q + A_'(rho z - y)
    auto publish_rhs = [&]() {
      const double e0 = kap[0] * (z[0] - u[0]), e1 = kap[1] * (z[1] - u[1]), e2 = kap[2] * (z[2] - u[2]),
                   e3 = kap[3] * (z[3] - u[3]), e4 = kap[4] * (z[4] - u[4]);
      rh[0] = sigma * x[0] - qb[0] + (e0 + e1);
      rh[1] = sigma * x[1] - qb[1] + (e2 + e3);
      rh[2] = sigma * x[2] - qb[2] + (e4 + (tzx * (e0 - e1) + tzy * (e2 - e3)));
      if (isleg) { sm.rhs[j0] = rh[0]; sm.rhs[j0 + 1] = rh[1]; sm.rhs[j0 + 2] = rh[2]; }
    };

    // per-lane constants of the recursions (axis role)
    const double fa = (c == 0) ? cyaw : (c == 1) ? -syaw : 0.0, fb = (c == 0) ? syaw : (c == 1) ? cyaw : 0.0;
    const double ba = (c == 0) ? cyaw : (c == 1) ? syaw : 0.0, bb = (c == 0) ? -syaw : (c == 1) ? cyaw : 0.0;
    const double fo = (c < 2) ? 0.0 : 1.0;
    const bool glast = (k == klast);
    // forward: reads X_k (group boundary for team 0), writes X_k+1; backward: reads p_k+1, writes p_k
    const double* const xin = (team == 0) ? &sm.Xb[12 * warp] : &sm.Xv[12 * k];
    double* const xout = glast ? &sm.Xb[12 * (warp + 1)] : &sm.Xv[12 * (k + 1)];
    const double* const pin = glast ? &sm.Pb[12 * (warp + 1)] : &sm.pv[12 * (k + 1)];
    double* const pout = (team == 0) ? &sm.Pb[12 * warp] : &sm.pv[12 * k];
    const double* const xin1 = (team == 0) ? sm.zero12 : xin;   // first sweep: zero boundaries
    const double* const pin1 = glast ? sm.zero12 : pin;
    const bool pb_zero = (warp == nG - 1);                      // p_H = 0 is this group's true boundary
    const double* const pin3 = (glast && pb_zero) ? sm.zero12 : pin;

    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    bool need_factor = true;

    for (;;) {
      if (need_factor) {
        need_factor = false;
        // ================= K3b: factorisation =================
        // Delta_leg = diag(c D^2 r2 + sigma) + A_leg' rho A_leg  (3 x 3, zero xy entry), inverse by cofactors
        {
          const double sa = kap[0] + kap[1], sb = kap[2] + kap[3];
          const double dxx = cs * D[0] * D[0] * r2x + sigma + sa, dyy = cs * D[1] * D[1] * r2y + sigma + sb;
          const double dzz = cs * D[2] * D[2] * r2z + sigma + kap[4] + tzx * tzx * sa + tzy * tzy * sb;
          const double dxz = tzx * (kap[0] - kap[1]), dyz = tzy * (kap[2] - kap[3]);
          const double m00 = dyy * dzz - dyz * dyz, m11 = dxx * dzz - dxz * dxz, m22 = dxx * dyy;
          const double idet = 1.0 / (dxx * m00 - dxz * dxz * dyy);
          di[0] = m00 * idet; di[1] = dxz * dyz * idet; di[2] = -dyy * dxz * idet;
          di[3] = m11 * idet; di[4] = -dxx * dyz * idet; di[5] = m22 * idet;
        }
        // G_ (into Fk) and M1 = G_ Delta^-1 (into Mt): the leg's three columns.  G_ = dt B6c D: the step's
        // velocity increment per unit of (scaled) force
        if (isleg) {
          const double dm[3][3] = {{di[0], di[1], di[2]}, {di[1], di[3], di[4]}, {di[2], di[4], di[5]}};
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double gq[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const double b6 = (cc < 3) ? top[cc < 3 ? cc : 0][q] : ((cc - 3 == q) ? inv_m : 0.0);
              gq[q] = dt * D[q] * b6;
              sm.Fk[k][cc * kMS + 3 * lg + q] = gq[q];
            }
#pragma unroll
            for (int q = 0; q < 3; ++q)
              sm.Mt[k][cc * kMS + 3 * lg + q] = gq[0] * dm[0][q] + gq[1] * dm[1][q] + gq[2] * dm[2][q];
          }
        }
        __syncwarp();
        // N = M1 G_' (row c per axis lane)
        if (isax) {
          double mr[12];
#pragma unroll
          for (int h2 = 0; h2 < 6; ++h2) {
            const double2 v = *reinterpret_cast<const double2*>(&sm.Mt[k][c * kMS + 2 * h2]);
            mr[2 * h2] = v.x; mr[2 * h2 + 1] = v.y;
          }
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            double s = 0.0;
#pragma unroll
            for (int h2 = 0; h2 < 6; ++h2) {
              const double2 v = *reinterpret_cast<const double2*>(&sm.Fk[k][d * kMS + 2 * h2]);
              s = fma(mr[2 * h2], v.x, s);
              s = fma(mr[2 * h2 + 1], v.y, s);
            }
            sm.Nk[k][6 * c + d] = s;
          }
        }
        __syncwarp();
        // Cholesky N_k = L L' (one lane per step; a non-positive pivot zeroes its column)
        if (on && t8 == 0) {
          double Lm[6][6];
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Lm[a][b] = sm.Nk[k][6 * a + b];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double d = Lm[cc][cc];
#pragma unroll
            for (int q = 0; q < cc; ++q) d -= Lm[cc][q] * Lm[cc][q];
            const bool ok = d > 0.0;
            const double ld = ok ? sqrt(d) : 0.0, li = ok ? 1.0 / ld : 0.0;
            Lm[cc][cc] = ld;
            sm.Li[k][cc] = li;
#pragma unroll
            for (int a = cc + 1; a < 6; ++a) {
              double s = Lm[a][cc];
#pragma unroll
              for (int q = 0; q < cc; ++q) s -= Lm[a][q] * Lm[cc][q];
              Lm[a][cc] = s * li;
            }
          }
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) sm.Lk[k][6 * a + b] = (b <= a) ? Lm[a][b] : 0.0;
        }
        __syncwarp();
        // M~ = L^-T L^-1 M1: the leg's three columns, forward then backward substitution
        if (isleg) {
          const double* Lp = sm.Lk[k];
          const double* li = sm.Li[k];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            double y6[6];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
              double s = sm.Mt[k][cc * kMS + 3 * lg + q];
#pragma unroll
              for (int b = 0; b < cc; ++b) s -= Lp[6 * cc + b] * y6[b];
              y6[cc] = s * li[cc];
            }
#pragma unroll
            for (int cc = 5; cc >= 0; --cc) {
              double s = y6[cc];
#pragma unroll
              for (int b = cc + 1; b < 6; ++b) s -= Lp[6 * b + cc] * y6[b];
              y6[cc] = s * li[cc];
            }
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) sm.Mt[k][cc * kMS + 3 * lg + q] = y6[cc];
          }
        }
        __syncthreads();
        // Riccati recursion on warp 0 (scratch in scr): Pi 144 | Y 144 | G 144 | T1 36 | Mm 36 | T2 36
        if (warp == 0) {
          double* const Pi = sm.scr;
          double* const Y = sm.scr + 144;
          double* const G = sm.scr + 288;
          double* const T1 = sm.scr + 432;
          double* const Mm = sm.scr + 468;
          double* const T2 = sm.scr + 504;
          for (int i = lane; i < 144; i += 32) Pi[i] = (i / 12 == i % 12) ? cs * bp.Qd[i / 12] : 0.0;
          __syncwarp();
#pragma unroll 1
          for (int ks = H - 1; ks >= 0; --ks) {
            const double* Lp = sm.Lk[ks];
            // Y = Pi A:  columns 0-5 unchanged, column 6 + j gains dt (Pi[:, 0:6] Rt)[:, j]
            for (int e = lane; e < 144; e += 32) {
              const int i = e / 12, j = e - 12 * i;
              double v = Pi[e];
              if (j >= 6) {
                const int q = j - 6;
                const double add = (q == 0) ? (cyaw * Pi[12 * i] - syaw * Pi[12 * i + 1])
                                 : (q == 1) ? (syaw * Pi[12 * i] + cyaw * Pi[12 * i + 1])
                                            : Pi[12 * i + q];
                v = fma(dt, add, v);
              }
              Y[e] = v;
            }
            // T1 = Pi_vv L
            for (int e = lane; e < 36; e += 32) {
              const int a = e / 6, b = e - 6 * a;
              double s = 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) s = fma(Pi[12 * (6 + a) + 6 + q], Lp[6 * q + b], s);
              T1[e] = s;
            }
            __syncwarp();
            // G = A' Y: rows 0-5 unchanged, row 6 + i gains dt (Rt' Y[0:6, :])[i, :]
            for (int e = lane; e < 144; e += 32) {
              const int i = e / 12, j = e - 12 * i;
              double v = Y[e];
              if (i >= 6) {
                const int q = i - 6;
                const double add = (q == 0) ? (cyaw * Y[j] - syaw * Y[12 + j])
                                 : (q == 1) ? (syaw * Y[j] + cyaw * Y[12 + j])
                                            : Y[12 * q + j];
                v = fma(dt, add, v);
              }
              G[e] = v;
            }
            // Mm = I + L' T1
            for (int e = lane; e < 36; e += 32) {
              const int a = e / 6, b = e - 6 * a;
              double s = (a == b) ? 1.0 : 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) s = fma(Lp[6 * q + a], T1[6 * q + b], s);
              Mm[e] = s;
            }
            __syncwarp();
            // Mm^-1 by Gauss-Jordan, lane r < 6 holds row r, the pivot row travels by shuffle (pivots >= 1)
            {
              const int r = lane < 6 ? lane : 0;
              double row[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) row[j] = Mm[6 * r + j];
#pragma unroll
              for (int pv_ = 0; pv_ < 6; ++pv_) {
                double rowp[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) rowp[j] = __shfl_sync(0xffffffffu, row[j], pv_);
                const double d = 1.0 / rowp[pv_];
                const bool piv = (lane == pv_);
                const double f = row[pv_] * d;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                  const double upd = (j == pv_) ? -f : row[j] - f * rowp[j];
                  const double prw = (j == pv_) ? d : rowp[j] * d;
                  row[j] = piv ? prw : upd;
                }
              }
              __syncwarp();
              if (lane < 6) {
#pragma unroll
                for (int j = 0; j < 6; ++j) Mm[6 * lane + j] = row[j];
              }
            }
            __syncwarp();
            // T2 = L Mm^-1
            for (int e = lane; e < 36; e += 32) {
              const int a = e / 6, b = e - 6 * a;
              double s = 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) s = fma(Lp[6 * a + q], Mm[6 * q + b], s);
              T2[e] = s;
            }
            __syncwarp();
            // Z = T2 L'  (symmetric: the upper triangle is computed and mirrored)
            for (int e = lane; e < 36; e += 32) {
              const int a = e / 6, b = e - 6 * a;
              if (a <= b) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) s = fma(T2[6 * a + q], Lp[6 * b + q], s);
                sm.Zk[ks][6 * a + b] = s;
                sm.Zk[ks][6 * b + a] = s;
              }
            }
            __syncwarp();
            // F = -Z U,  U = Y[6:12, :]
            for (int e = lane; e < 72; e += 32) {
              const int a = e / 12, j = e - 12 * a;
              double s = 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) s = fma(sm.Zk[ks][6 * a + q], Y[12 * (6 + q) + j], s);
              sm.Fk[ks][a * kMS + j] = -s;
            }
            __syncwarp();
            // Pi <- cQ + G + U' F  (upper triangle, mirrored)
            for (int e = lane; e < 78; e += 32) {
              // e -> (i, j), i <= j, row by row of the upper triangle
              int i = 0, rem = e;
              while (rem >= 12 - i) { rem -= 12 - i; ++i; }
              const int j = i + rem;
              double s = 0.5 * (G[12 * i + j] + G[12 * j + i]);
#pragma unroll
              for (int q = 0; q < 6; ++q) s = fma(Y[12 * (6 + q) + i], sm.Fk[ks][q * kMS + j], s);
              if (i == j) s += cs * bp.Qd[i];
              Pi[12 * i + j] = s;
              Pi[12 * j + i] = s;
            }
            __syncwarp();
          }
        }
        __syncthreads();
        // group transitions Phi_j = Acl_(4j+3) ... Acl_4j, j = 1 .. nG-2: one thread per column
        if (tid < (nG - 2) * 12) {
          const int gj = 1 + tid / 12, col = tid % 12;
          double T[12], Tn[6];
#pragma unroll
          for (int i = 0; i < 12; ++i) T[i] = (i == col) ? 1.0 : 0.0;
#pragma unroll 1
          for (int s_ = 0; s_ < 4; ++s_) {
            const double* Fp = sm.Fk[4 * gj + s_];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              double acc = T[6 + a];
#pragma unroll
              for (int i = 0; i < 12; ++i) acc = fma(Fp[a * kMS + i], T[i], acc);
              Tn[a] = acc;
            }
            T[0] = fma(dt, cyaw * T[6] + syaw * T[7], T[0]);
            T[1] = fma(dt, -syaw * T[6] + cyaw * T[7], T[1]);
            T[2] = fma(dt, T[8], T[2]);
            T[3] = fma(dt, T[9], T[3]);
            T[4] = fma(dt, T[10], T[4]);
            T[5] = fma(dt, T[11], T[5]);
#pragma unroll
            for (int a = 0; a < 6; ++a) T[6 + a] = Tn[a];
          }
#pragma unroll
          for (int r = 0; r < 12; ++r) sm.Phi[gj - 1][r * kMS + col] = T[r];
        }
        publish_rhs();
        __syncthreads();
      }
      int run = until_check < until_adapt ? until_check : until_adapt;
      run = run < sp.max_iter - iter ? run : sp.max_iter - iter;
#pragma unroll 1
      for (int q_ = 0; q_ < run; ++q_) {
        // ---- u = M~ r (axis), a = Delta^-1 r (leg) ----
        __syncwarp();
        double uc = 0.0;
        {
          const double2* mp = reinterpret_cast<const double2*>(&sm.Mt[k][c * kMS]);
          const double2* rp = reinterpret_cast<const double2*>(&sm.rhs[12 * k]);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int h2 = 0; h2 < 6; ++h2) {
            const double2 mv = mp[h2], r = rp[h2];
            s0 = fma(mv.x, r.x, s0);
            s1 = fma(mv.y, r.y, s1);
          }
          uc = s0 + s1;
          if (isax) sm.uv[6 * k + c] = uc;
        }
        const double a0 = di[0] * rh[0] + di[1] * rh[1] + di[2] * rh[2];
        const double a1 = di[1] * rh[0] + di[3] * rh[1] + di[4] * rh[2];
        const double a2 = di[2] * rh[0] + di[4] * rh[1] + di[5] * rh[2];
        __syncwarp();
        // (N u)_c and t = -F' u (the addend of the backward recursion); F columns c and 6 + c stay in registers
        double u6[6], fc[12], nuc, tpos, tvel;
        {
          const double2* up = reinterpret_cast<const double2*>(&sm.uv[6 * k]);
          const double2 v0 = up[0], v1 = up[1], v2 = up[2];
          u6[0] = v0.x; u6[1] = v0.y; u6[2] = v1.x; u6[3] = v1.y; u6[4] = v2.x; u6[5] = v2.y;
          const double* np = &sm.Nk[k][6 * c];
          const double* fp = &sm.Fk[k][c];
          double s = 0.0, sp_ = 0.0, sv = 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            fc[d] = fp[d * kMS];
            fc[6 + d] = fp[d * kMS + 6];
            s = fma(np[d], u6[d], s);
            sp_ = fma(fc[d], u6[d], sp_);
            sv = fma(fc[6 + d], u6[d], sv);
          }
          nuc = s; tpos = -sp_; tvel = -sv;
        }
        // ---- backward recursion  p_k = Acl_k' p_k+1 + t_k ----
        auto bstep = [&](const double* src, double* dst, bool wr) {
          const double2* pp = reinterpret_cast<const double2*>(src);
          const double2 p01 = pp[0], v01 = pp[3], v23 = pp[4], v45 = pp[5];
          const double own_pos = src[c], own_vel = src[6 + c];
          double s0 = tpos, s1 = 0.0, s2 = tvel, s3 = 0.0;
          s0 = fma(fc[0], v01.x, s0); s1 = fma(fc[1], v01.y, s1);
          s2 = fma(fc[6], v01.x, s2); s3 = fma(fc[7], v01.y, s3);
          s0 = fma(fc[2], v23.x, s0); s1 = fma(fc[3], v23.y, s1);
          s2 = fma(fc[8], v23.x, s2); s3 = fma(fc[9], v23.y, s3);
          s0 = fma(fc[4], v45.x, s0); s1 = fma(fc[5], v45.y, s1);
          s2 = fma(fc[10], v45.x, s2); s3 = fma(fc[11], v45.y, s3);
          const double rot = fma(ba, p01.x, fma(bb, p01.y, fo * own_pos));
          const double np_ = own_pos + (s0 + s1);
          const double nv_ = fma(dt, rot, own_vel) + (s2 + s3);
          if (wr) { dst[c] = np_; dst[6 + c] = nv_; }
        };
#pragma unroll
        for (int s_ = 3; s_ >= 0; --s_) {
          if (team == s_) bstep(pin1, pout, isax);
          __syncwarp();
        }
        __syncthreads();
        if (warp == 0) {
          const int i = lane < 12 ? lane : 0;
#pragma unroll 1
          for (int gj = nG - 2; gj >= 1; --gj) {
            const double2* pp = reinterpret_cast<const double2*>(&sm.Pb[12 * (gj + 1)]);
            const double* ph = &sm.Phi[gj - 1][i];
            double s0 = sm.Pb[12 * gj + i], s1 = 0.0;
#pragma unroll
            for (int h2 = 0; h2 < 6; ++h2) {
              const double2 v = pp[h2];
              s0 = fma(ph[(2 * h2) * kMS], v.x, s0);
              s1 = fma(ph[(2 * h2 + 1) * kMS], v.y, s1);
            }
            if (lane < 12) sm.Pb[12 * gj + i] = s0 + s1;
            __syncwarp();
          }
        }
        __syncthreads();
#pragma unroll
        for (int s_ = 3; s_ >= 1; --s_) {
          if (team == s_) bstep(pin3, pout, isax);
          __syncwarp();
        }
        // ---- e = p_k+1,vel - u,  b = -Z e ----
        const double ec = pin3[6 + c] - uc;
        if (isax) sm.ev[6 * k + c] = ec;
        __syncwarp();
        double bfw;
        {
          const double2* ep = reinterpret_cast<const double2*>(&sm.ev[6 * k]);
          const double2 e01 = ep[0], e23 = ep[1], e45 = ep[2];
          const double* zp = &sm.Zk[k][6 * c];
          double s0 = zp[0] * e01.x, s1 = zp[1] * e01.y;
          s0 = fma(zp[2], e23.x, s0); s1 = fma(zp[3], e23.y, s1);
          s0 = fma(zp[4], e45.x, s0); s1 = fma(zp[5], e45.y, s1);
          bfw = -(s0 + s1);
        }
        // ---- forward recursion  dlt_k = F_k X_k + b_k,  X_k+1 = A X_k + [0; dlt_k] ----
        double fr[12], dl = 0.0;
        {
          const double2* fp = reinterpret_cast<const double2*>(&sm.Fk[k][c * kMS]);
#pragma unroll
          for (int h2 = 0; h2 < 6; ++h2) { const double2 v = fp[h2]; fr[2 * h2] = v.x; fr[2 * h2 + 1] = v.y; }
        }
        auto fstep = [&](const double* src, double* dst, bool wr) {
          const double2* xp = reinterpret_cast<const double2*>(src);
          const double own_pos = src[c], own_vel = src[6 + c];
          double s0 = bfw, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          double2 w01;
#pragma unroll
          for (int h2 = 0; h2 < 6; h2 += 2) {
            const double2 va = xp[h2], vb = xp[h2 + 1];
            s0 = fma(fr[2 * h2], va.x, s0); s1 = fma(fr[2 * h2 + 1], va.y, s1);
            s2 = fma(fr[2 * h2 + 2], vb.x, s2); s3 = fma(fr[2 * h2 + 3], vb.y, s3);
          }
          w01 = xp[3];
          dl = (s0 + s1) + (s2 + s3);
          const double rot = fma(fa, w01.x, fma(fb, w01.y, fo * own_vel));
          if (wr) { dst[c] = fma(dt, rot, own_pos); dst[6 + c] = own_vel + dl; }
        };
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_) {
          if (team == s_) fstep(xin1, xout, isax);
          __syncwarp();
        }
        __syncthreads();
        if (warp == 0) {
          const int i = lane < 12 ? lane : 0;
#pragma unroll 1
          for (int gj = 1; gj <= nG - 2; ++gj) {
            const double2* xp = reinterpret_cast<const double2*>(&sm.Xb[12 * gj]);
            const double2* ph = reinterpret_cast<const double2*>(&sm.Phi[gj - 1][i * kMS]);
            double s0 = sm.Xb[12 * (gj + 1) + i], s1 = 0.0;
#pragma unroll
            for (int h2 = 0; h2 < 6; ++h2) {
              const double2 v = xp[h2], f = ph[h2];
              s0 = fma(f.x, v.x, s0);
              s1 = fma(f.y, v.y, s1);
            }
            if (lane < 12) sm.Xb[12 * (gj + 1) + i] = s0 + s1;
            __syncwarp();
          }
        }
        __syncthreads();
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_) {
          if (team == s_) fstep(xin, xout, isax && !glast);
          __syncwarp();
        }
        // ---- h = N u - dlt (axis);  x~ = a - M~' h, row updates, next rhs (leg) ----
        if (isax) sm.ev[6 * k + c] = nuc - dl;
        __syncwarp();
        {
          const double2* hp = reinterpret_cast<const double2*>(&sm.ev[6 * k]);
          const double2 h01 = hp[0], h23 = hp[1], h45 = hp[2];
          const double h6[6] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y};
          const double* mp = &sm.Mt[k][3 * lg];
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            s0 = fma(mp[cc * kMS], h6[cc], s0);
            s1 = fma(mp[cc * kMS + 1], h6[cc], s1);
            s2 = fma(mp[cc * kMS + 2], h6[cc], s2);
          }
          const double xtx = a0 - s0, xty = a1 - s1, xtz = a2 - s2;
          x[0] = alpha * xtx + (1.0 - alpha) * x[0];
          x[1] = alpha * xty + (1.0 - alpha) * x[1];
          x[2] = alpha * xtz + (1.0 - alpha) * x[2];
          const double zt[5] = {fma(tzx, xtz, xtx), fma(-tzx, xtz, xtx), fma(tzy, xtz, xty), fma(-tzy, xtz, xty), xtz};
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const double zr = alpha * zt[i] + (1.0 - alpha) * z[i];
            double zn = zr + u[i];
            if (i == 0 || i == 2) zn = (zn < 0.0) ? 0.0 : zn;
            else if (i == 1 || i == 3) zn = (zn > 0.0) ? 0.0 : zn;
            else { zn = (zn < lo4) ? lo4 : zn; zn = (zn > hi4) ? hi4 : zn; }
            u[i] = u[i] + (zr - zn);
            z[i] = zn;
          }
          publish_rhs();
        }
      }
      __syncwarp();
      iter += run;
      until_check -= run;
      until_adapt -= run;
      const bool can_check = (until_check == 0);
      const bool can_adapt = (until_adapt == 0);
      if (can_check) until_check = sp.check_termination;
      if (can_adapt) until_adapt = sp.adaptive_rho_interval;
      const bool last = (iter == sp.max_iter);

      // ---- residuals (auxil.c compute_pri_res / compute_dua_res / tolerances) ----
      const double cinv = sm.scal[1];
      double* const xD = sm.scr;            // D x            (n)
      double* const vt = sm.scr + n;        // G D x          (6H)
      double* const vo = sm.scr + n + 6 * H;  // S G D x      (6H)
      if (isleg) { xD[j0] = D[0] * x[0]; xD[j0 + 1] = D[1] * x[1]; xD[j0 + 2] = D[2] * x[2]; }
      __syncwarp();
      if (isax) {
        // u = B6c (D x) of the step
        const double* xd = &xD[12 * k];
        double s = 0.0;
        if (c < 3) {
#pragma unroll
          for (int i = 0; i < 12; ++i) s = fma(fma(-(double)k, sm.dT[3 * c + i % 3], sm.B6t[c][i]), xd[i], s);
        } else {
          s = inv_m * ((xd[c - 3] + xd[c]) + (xd[c + 3] + xd[c + 6]));
        }
        vt[6 * k + c] = s;
      }
      __syncthreads();
      if (isax) {
        // w = S u = D1 (alpha u) + D2 (beta u), one lane per wrench component
        double sa = 0.0, sb0 = 0.0, sb1 = 0.0;
        const int o0 = (c < 2) ? 0 : c, o1 = (c < 2) ? 1 : c;
        for (int l = 0; l < H; ++l) {
          const int mxk = k > l ? k : l;
          const double a = (double)(H - mxk) * dt2, b = dt4 * (double)sm.be[H * k + l];
          sa = fma(a, vt[6 * l + c], sa);
          sb0 = fma(b, vt[6 * l + o0], sb0);
          sb1 = fma(b, vt[6 * l + o1], sb1);
        }
        double w;
        if (c == 0) w = bp.Qd[6] * sa + th00 * sb0 + th01 * sb1;
        else if (c == 1) w = bp.Qd[7] * sa + th01 * sb0 + th11 * sb1;
        else if (c == 2) w = bp.Qd[8] * sa + th22 * sb0;
        else w = bp.Qd[6 + c] * sa + bp.Qd[c] * sb0;
        vo[6 * k + c] = w;
      }
      __syncwarp();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      if (isleg) {
        const double Ah[5] = {fma(tzx, x[2], x[0]), fma(-tzx, x[2], x[0]), fma(tzy, x[2], x[1]), fma(-tzy, x[2], x[1]), x[2]};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const double Dn = (i < 2) ? D[0] : (i < 4) ? D[1] : D[2];
          const double rp = Ah[i] - z[i];
          v[0] = fmax(v[0], fabs(cca[i] * rp));   // scaled primal residual
          v[1] = fmax(v[1], Dn * fabs(rp));       // unscaled: E^-1 cca = D
          v[2] = fmax(v[2], fabs(Dn * z[i]));
          v[3] = fmax(v[3], fabs(Dn * Ah[i]));
          v[4] = fmax(v[4], fabs(cca[i] * z[i]));
          v[5] = fmax(v[5], fabs(cca[i] * Ah[i]));
        }
        // P_ x = c D (R2 D x + G' w) ; A_' y
        const double* wv = &vo[6 * k];
        const double f0 = kap[0] * u[0], f1 = kap[1] * u[1], f2 = kap[2] * u[2], f3 = kap[3] * u[3], f4 = kap[4] * u[4];
        const double Aty[3] = {f0 + f1, f2 + f3, f4 + (tzx * (f0 - f1) + tzy * (f2 - f3))};
        const double r2v[3] = {r2x, r2y, r2z};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const double gtw = top[0][q] * wv[0] + top[1][q] * wv[1] + top[2][q] * wv[2] + inv_m * wv[3 + q];
          const double Px = cs * D[q] * (r2v[q] * (D[q] * x[q]) + gtw);
          const double Dinv = 1.0 / D[q];
          const double rd = Px + qb[q] + Aty[q];
          v[6] = fmax(v[6], fabs(rd));
          v[7] = fmax(v[7], fabs(Dinv * rd));
          v[8] = fmax(v[8], fmax(fmax(fabs(Dinv * qb[q]), fabs(Dinv * Aty[q])), fabs(Dinv * Px)));
          v[9] = fmax(v[9], fmax(fmax(fabs(qb[q]), fabs(Aty[q])), fabs(Px)));
        }
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] = fmax(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) sm.red[warp * 16 + i] = v[i];
      }
      __syncthreads();
      if (tid == 0) {
        double mres[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          double t = sm.red[i];
#pragma unroll
          for (int w = 1; w < kWrcWarps; ++w) t = fmax(t, sm.red[w * 16 + i]);
          mres[i] = t;
        }
        const double pri = mres[1], dua = cinv * mres[7];
        const double eps_pri = sp.eps_abs + sp.eps_rel * fmax(mres[2], mres[3]);
        const double eps_dua = sp.eps_abs + sp.eps_rel * cinv * mres[8];
        sm.scal[4] = pri;
        int done = 0, refactor = 0;
        if ((can_check || last) && pri < eps_pri && dua < eps_dua) {
          done = 1;
          sm.flags[1] = MPC_STATUS_SOLVED;
        } else if (last) {
          done = 1;
          sm.flags[1] = (pri < 10.0 * eps_pri && dua < 10.0 * eps_dua) ? 2 : MPC_STATUS_MAX_ITER_REACHED;
        } else if (can_adapt) {
          const double rho_c = sm.scal[2];
          const double pn = mres[0] / (fmax(mres[4], mres[5]) + 1e-10);
          const double dn = mres[6] / (mres[9] + 1e-10);
          double rho_new = rho_c * sqrt(pn / (dn + 1e-10));
          rho_new = fmin(fmax(rho_new, 1e-6), 1e6);
          if (rho_new > rho_c * sp.adaptive_rho_tolerance || rho_new < rho_c / sp.adaptive_rho_tolerance) {
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
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const double rn = rho_of(ct[i], rho);
          u[i] *= rv[i] / rn;   // y stays, uh = y / (rho cca) follows the new rho
          rv[i] = rn;
          kap[i] = rn * cca[i] * cca[i];
        }
        need_factor = true;  // the factorisation ends by rebuilding the right-hand side with the new rho vector
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    if (kWarm) {
      // keep the solver alive for the next tick -- unless this solve went wrong (slot marked dead: next tick
      // is an initSolver)
      bool fin = true;
#pragma unroll
      for (int q = 0; q < 3; ++q) fin = fin && isfinite(x[q]);
#pragma unroll
      for (int i = 0; i < 5; ++i) fin = fin && isfinite(z[i]) && isfinite(u[i]);
      const int all_ok = __syncthreads_and(fin || !isleg);
      if (isleg) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          ws[kWX + j0 + q] = x[q];
          ws[kWQ + j0 + q] = q0v[q];
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          ws[kWZ + r0 + i] = cca[i] * z[i];            // OSQP's scaled z, y
          ws[kWY + r0 + i] = rv[i] * cca[i] * u[i];
        }
      }
      if (tid == 0) {
        ws[kWRho] = sm.scal[2];
        ws[kWLive] = (all_ok && isfinite(sm.scal[2])) ? 1.0 : 0.0;
      }
    }
    const double f0 = D[0] * x[0], f1 = D[1] * x[1], f2 = D[2] * x[2];
    if (x_all != nullptr && isleg) {
      x_all[size_t(p) * n + j0] = (float)f0;
      x_all[size_t(p) * n + j0 + 1] = (float)f1;
      x_all[size_t(p) * n + j0 + 2] = (float)f2;
    }
    if (warp == 0) {
      // step 0 is team 0 of warp 0: lanes 0..3 hold the legs
      const bool first = lane < 4;
      double gb[3] = {0.0, 0.0, 0.0};
      int nanbits = 0;
      if (first) {
        const float* R = st + kOffRot;  // R' f (A1RobotControl.cpp:558-561)
        const bool bad = isnan(f0) || isnan(f1) || isnan(f2);  // NaN guard (:559)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const double gv = (double)R[q] * f0 + (double)R[3 + q] * f1 + (double)R[6 + q] * f2;
          gb[q] = bad ? 0.0 : gv;
          results[p].grf[3 * lane + q] = (float)gb[q];
        }
        if (tin != nullptr) {
          const MpcTorqueIn& t = tin[p];
          const bool contact = st[kOffContacts + lane] != 0.0f;
          double tau[3];
          leg_torque(t.j_foot + 9 * lane, contact, gb[0], gb[1], gb[2], t.foot_forces_kin + 3 * lane, t.km_foot,
                     t.torques_gravity + 3 * lane, tau);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const bool tn = isnan(tau[q]);
            nanbits |= tn ? (1 << (3 * lane + q)) : 0;
            tout[p].joint_torques[3 * lane + q] = tn ? 0.0f : (float)tau[q];
          }
        }
      }
      if (tin != nullptr) {
        nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 1);
        nanbits |= __shfl_xor_sync(0xffffffffu, nanbits, 2);
        if (lane == 0) tout[p].nan_mask = (int32_t)(nanbits & 0xfff);
      }
    }
    if (tid == 0) {
      results[p].status = status;
      results[p].iters = iter;
      results[p].rho_updates = rho_updates;
      results[p].pri_res = (float)pri_res_out;
    }
  }
}

}  // namespace mpcb200
