// Generic-horizon kernels (template on H; used for H = 30, BASELINE config 4).
//
// Same algorithms and the same arithmetic (f64) as the H = 10 kernels, but a 360-variable
// problem does not fit one SM: K^-1 alone is 1.04 MB in f64 (registers 256 KB, smem 227 KB).
// This first long-horizon path keeps B_qp (build) and -K^-1 (solve) in a PER-CTA workspace in
// global memory -- 148 persistent CTAs x 1.04 MB, mostly L2-resident (126 MB) -- and streams it
// with coalesced warp-per-row passes; vectors live in shared memory.  The shipped long-horizon path
// is the Riccati-structured solver (riccati_kernel.cuh, 1/40 of the flops and no 1 MB inverse); this
// dense one stays as structured_solver = 2, an independent second implementation for the tests.
//
//   gen_build_kernel<H>   K0+K1+K2  (ConvexMpc.cpp:110-245, A1RobotControl.cpp:452-518)
//   gen_solve_kernel<H>   K3+K4+K5  (OSQP 0.6.x as driven by A1RobotControl.cpp:522-561)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_kernels.cuh"

namespace mpcb200 {

constexpr int kGenBuildThreads = 256;
constexpr int kGenSolveThreads = 512;
constexpr int kGenSolveWarps = kGenSolveThreads / 32;
constexpr int kTile = 60;    // Hessian output tile (n = 12 H must be a multiple of 60: H % 5 == 0)
constexpr int kKChunk = 26;  // k rows per staged chunk (s = 13 H must be a multiple of 26: H even)

template <int H>
struct GenBuildSmem {
  double Apow[(H + 1) * 169];
  alignas(16) double Bd[H * 156];
  double S[H * 169];   // first C_m = (A^m)' Q A^m, then their running sums
  double T[H * 156];   // B_j' S_j
  double xref[13 * H];
  double tmp[13 * H];
  double g[13 * H];
  double x0[16];
  alignas(16) double scratch[kGenBuildThreads / 32][312];  // per warp: U (13 x 12, padded to 160) and a 12 x 12 block
  float st[48];
  int contacts[4 * H];  // per step and leg (replicated unless the horizon is gait aware)
};

template <int H>
__global__ void __launch_bounds__(kGenBuildThreads, 1)
gen_build_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait, ModelIn model, int num, double* __restrict__ model_out,
                 double* __restrict__ P_out,
                 double* __restrict__ q_out, float* __restrict__ l_out, float* __restrict__ u_out,
                 double* __restrict__ workspace, const __grid_constant__ BuildParams bp) {
  constexpr int n = 12 * H, s = 13 * H, m = 20 * H;
  static_assert(n % kTile == 0 && s % kKChunk == 0, "horizon must be a multiple of 10");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GenBuildSmem<H>& sm = *reinterpret_cast<GenBuildSmem<H>*>(smem_raw);
  const int tid = threadIdx.x;
  (void)workspace;  // the build no longer needs scratch in global memory (B_qp is never formed)

  for (int p = blockIdx.x; p < num; p += gridDim.x) {
    __syncthreads();
    if (states != nullptr) {
      if (tid < 48) sm.st[tid] = reinterpret_cast<const float*>(states + p)[tid];
      __syncthreads();
      const float* st = sm.st;
      if (tid < 169) {  // A_d = I + dt A_c (ConvexMpc.cpp:110-130, :149-150)
        const int rr = tid / 13, cc = tid % 13;
        double sy, cy;
        sincos((double)st[kOffEuler + 2], &sy, &cy);
        double a = 0.0;
        if (rr == 0 && cc == 6) a = cy;
        if (rr == 0 && cc == 7) a = sy;
        if (rr == 1 && cc == 6) a = -sy;
        if (rr == 1 && cc == 7) a = cy;
        if (rr == 2 && cc == 8) a = 1.0;
        if (rr >= 3 && rr <= 5 && cc == rr + 6) a = 1.0;
        if (rr == 11 && cc == 12) a = 1.0;
        // exact discretisation: A_c^3 = 0 and A_c^2 has the single entry (5, 12) = 1 (see qp_build_kernel)
        if (bp.exact_discretization && rr == 5 && cc == 12) a = 0.5 * bp.dt;
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = ((rr == cc) ? 1.0 : 0.0) + a * bp.dt;
      }
      for (int idx = tid; idx < H * 156; idx += kGenBuildThreads) sm.Bd[idx] = 0.0;
      if (tid >= 192 && tid < 192 + 13) sm.x0[tid - 192] = (tid - 192 < 12) ? (double)st[tid - 192] : -9.8;
      for (int idx = tid; idx < 4 * H; idx += kGenBuildThreads) {
        const int i = idx >> 2, leg = idx & 3;
        int c = st[kOffContacts + leg] != 0.0f;
        if (bp.gait_aware && i > 0) {
          // planned contact of step i from the gait counter (A1RobotControl.cpp:156-164)
          const float* gg = reinterpret_cast<const float*>(gait + p);
          const double cnt = fmod((double)gg[leg] + (double)i * (double)gg[10] * (double)gg[4 + leg], (double)gg[8]);
          c = cnt <= (double)gg[9];
        }
        sm.contacts[idx] = c;
      }
      for (int i = tid; i < H; i += kGenBuildThreads) {  // mpc_states_d (A1RobotControl.cpp:470-488)
        const double R0 = st[kOffRot + 0], R1 = st[kOffRot + 1], R2 = st[kOffRot + 2];
        const double R3 = st[kOffRot + 3], R4 = st[kOffRot + 4], R5 = st[kOffRot + 5];
        const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
        const double vwx = R0 * vx + R1 * vy + R2 * vz, vwy = R3 * vx + R4 * vy + R5 * vz;
        double* d = &sm.xref[13 * i];
        d[0] = (double)st[kOffEulerD];
        d[1] = (double)st[kOffEulerD + 1];
        d[2] = (double)st[kOffEuler + 2] + (double)st[kOffAngVelD + 2] * bp.dt * (double)(i + 1);
        d[3] = (double)st[kOffPos] + vwx * bp.dt * (double)(i + 1);
        d[4] = (double)st[kOffPos + 1] + vwy * bp.dt * (double)(i + 1);
        d[5] = (double)st[kOffPosDz];
        d[6] = (double)st[kOffAngVelD];
        d[7] = (double)st[kOffAngVelD + 1];
        d[8] = (double)st[kOffAngVelD + 2];
        d[9] = vwx;
        d[10] = vwy;
        d[11] = 0.0;
        d[12] = -9.8;
      }
      __syncthreads();
      // B_d = dt B_c, one thread per leg (ConvexMpc.cpp:132-143, :151) -- per (step, leg) when the feet drift
      if (tid < (bp.foot_drift ? 4 * H : 4)) {
        const int leg = tid & 3, step = tid >> 2;
        double* Bs_ = &sm.Bd[step * 156];
        double R[9], T[9], Iw[9], Inv[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (double)st[kOffRot + i];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) a += R[3 * i + k] * bp.inertia[3 * k + j];
            T[3 * i + j] = a;
          }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) a += T[3 * i + k] * R[3 * j + k];
            Iw[3 * i + j] = a;
          }
        const double c00 = Iw[4] * Iw[8] - Iw[5] * Iw[7], c01 = Iw[5] * Iw[6] - Iw[3] * Iw[8],
                     c02 = Iw[3] * Iw[7] - Iw[4] * Iw[6];
        const double id = 1.0 / (Iw[0] * c00 + Iw[1] * c01 + Iw[2] * c02);
        Inv[0] = c00 * id; Inv[1] = (Iw[2] * Iw[7] - Iw[1] * Iw[8]) * id; Inv[2] = (Iw[1] * Iw[5] - Iw[2] * Iw[4]) * id;
        Inv[3] = c01 * id; Inv[4] = (Iw[0] * Iw[8] - Iw[2] * Iw[6]) * id; Inv[5] = (Iw[2] * Iw[3] - Iw[0] * Iw[5]) * id;
        Inv[6] = c02 * id; Inv[7] = (Iw[1] * Iw[6] - Iw[0] * Iw[7]) * id; Inv[8] = (Iw[0] * Iw[4] - Iw[1] * Iw[3]) * id;
        double fx = st[kOffFoot + 3 * leg], fy = st[kOffFoot + 3 * leg + 1], fz = st[kOffFoot + 3 * leg + 2];
        if (bp.foot_drift) {
          // the body moves on with the commanded world velocity, the stance feet stay: r_i = r_0 - i dt v_d
          const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
          const double kd = (double)step * bp.dt;
          fx -= kd * (R[0] * vx + R[1] * vy + R[2] * vz);
          fy -= kd * (R[3] * vx + R[4] * vy + R[5] * vz);
          fz -= kd * (R[6] * vx + R[7] * vy + R[8] * vz);
        }
        const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) a += Inv[3 * i + k] * sk[3 * k + j];
            Bs_[(6 + i) * 12 + 3 * leg + j] = a * bp.dt;
            Bs_[(9 + i) * 12 + 3 * leg + j] = (i == j) ? (1.0 / bp.mass) * bp.dt : 0.0;
          }
        if (bp.exact_discretization) {
          // B_d += dt^2/2 A_c B_c: euler rows <- Rz' (I^-1 [r]x), position rows <- I / m
          double sy, cy;
          sincos((double)st[kOffEuler + 2], &sy, &cy);
          const double hh = 0.5 * bp.dt;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double b6 = Bs_[6 * 12 + 3 * leg + j], b7 = Bs_[7 * 12 + 3 * leg + j], b8 = Bs_[8 * 12 + 3 * leg + j];
            Bs_[0 * 12 + 3 * leg + j] = hh * (cy * b6 + sy * b7);
            Bs_[1 * 12 + 3 * leg + j] = hh * (-sy * b6 + cy * b7);
            Bs_[2 * 12 + 3 * leg + j] = hh * b8;
#pragma unroll
            for (int i = 0; i < 3; ++i) Bs_[(3 + i) * 12 + 3 * leg + j] = hh * Bs_[(9 + i) * 12 + 3 * leg + j];
          }
        }
      }
      __syncthreads();
      if (!bp.foot_drift)
        for (int idx = tid; idx < (H - 1) * 156; idx += kGenBuildThreads) sm.Bd[156 + idx] = sm.Bd[idx % 156];
    } else {
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        sm.Apow[tid] = (rr == cc) ? 1.0 : 0.0;
        sm.Apow[169 + tid] = model.A_d[size_t(p) * 169 + tid];
      }
      for (int idx = tid; idx < H * 156; idx += kGenBuildThreads) sm.Bd[idx] = model.B_d_list[size_t(p) * H * 156 + idx];
      if (tid < 13) sm.x0[tid] = model.x0[size_t(p) * 13 + tid];
      for (int idx = tid; idx < s; idx += kGenBuildThreads) sm.xref[idx] = model.x_ref[size_t(p) * s + idx];
      for (int idx = tid; idx < 4 * H; idx += kGenBuildThreads) sm.contacts[idx] = model.contacts[size_t(p) * 4 + (idx & 3)] != 0;
    }
    if (model_out != nullptr) {
      // A_d and the B_d list for the structured (Riccati) solver
      double* mo = model_out + size_t(p) * (169 + H * 156);
      for (int idx = tid; idx < 169 + H * 156; idx += kGenBuildThreads)
        mo[idx] = (idx < 169) ? sm.Apow[169 + idx] : sm.Bd[idx - 169];
    }
    __syncthreads();
    // A_qp powers (ConvexMpc.cpp:185-191)
    for (int i = 1; i < H; ++i) {
      if (tid < 169) {
        const int rr = tid / 13, cc = tid % 13;
        const double* Ap = &sm.Apow[i * 169];
        const double* A1 = &sm.Apow[169];
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 13; ++k) a += Ap[rr * 13 + k] * A1[k * 13 + cc];
        sm.Apow[(i + 1) * 169 + tid] = a;
      }
      __syncthreads();
    }
    // Condensed Hessian and gradient WITHOUT B_qp, as in qp_build_kernel (mpc_kernels.cuh):
    //   block (j, l), j >= l:  T_j (A^(j-l) B_l),  T_j = B_j' S_j,  S_j = sum_{m <= H-1-j} (A^m)' Q A^m
    //   gradient block j:      B_j' g_j,  g_j = sum_{i >= j} (A^(i-j))' Q (A^(i+1) x0 - x_ref,i)
    for (int idx = tid; idx < H * 169; idx += kGenBuildThreads) {  // C_m = (A^m)' Q A^m
      const int mm = idx / 169, e = idx - 169 * mm, ia = e / 13, ib = e - 13 * ia;
      const double* Am = &sm.Apow[mm * 169];
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) a = fma(Am[k * 13 + ia] * bp.Qd[k], Am[k * 13 + ib], a);
      sm.S[idx] = a;
    }
    for (int idx = tid; idx < s; idx += kGenBuildThreads) {  // tmp_i = Q (A^(i+1) x0 - x_ref,i)
      const int i = idx / 13, rr = idx % 13;
      const double* Ai = &sm.Apow[(i + 1) * 169 + rr * 13];
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 13; ++k) a += Ai[k] * sm.x0[k];
      sm.tmp[idx] = bp.Qd[rr] * (a - sm.xref[idx]);
    }
    __syncthreads();
    if (tid < 169) {  // running sums in place: S[mm] = C_0 + .. + C_mm, so S_j = S[H-1-j]
      double acc = 0.0;
      for (int mm = 0; mm < H; ++mm) {
        acc += sm.S[mm * 169 + tid];
        sm.S[mm * 169 + tid] = acc;
      }
    }
    for (int idx = tid; idx < s; idx += kGenBuildThreads) {  // g_j
      const int j = idx / 13, ia = idx - 13 * j;
      double a = 0.0;
      for (int i = j; i < H; ++i) {
        const double* Am = &sm.Apow[(i - j) * 169];
#pragma unroll
        for (int k = 0; k < 13; ++k) a = fma(Am[k * 13 + ia], sm.tmp[13 * i + k], a);
      }
      sm.g[idx] = a;
    }
    __syncthreads();
    for (int idx = tid; idx < H * 156; idx += kGenBuildThreads) {  // T_j = B_j' S_j (12 x 13)
      const int j = idx / 156, e = idx - 156 * j, ia = e / 13, k = e - 13 * ia;
      const double* Bj = &sm.Bd[j * 156];
      const double* Sj = &sm.S[(H - 1 - j) * 169];
      double a = 0.0;
#pragma unroll
      for (int pp = 0; pp < 13; ++pp) a = fma(Bj[pp * 12 + ia], Sj[pp * 13 + k], a);
      sm.T[idx] = a;
    }
    for (int cidx = tid; cidx < n; cidx += kGenBuildThreads) {  // gradient = B_j' g_j
      const int j = cidx / 12, ia = cidx - 12 * j;
      const double* Bj = &sm.Bd[j * 156];
      double a = 0.0;
#pragma unroll
      for (int pp = 0; pp < 13; ++pp) a = fma(Bj[pp * 12 + ia], sm.g[13 * j + pp], a);
      q_out[size_t(p) * n + cidx] = a;
    }
    __syncthreads();
    // one WARP per block (j, l), l <= j: U = A^(j-l) B_l into the warp's scratch, then the 12 x 12
    // block T_j U; the block and its mirror leave as rows of 12 contiguous doubles
    {
      double* Pp = P_out + size_t(p) * n * n;
      const int warp = tid >> 5, lane = tid & 31;
      double* U = sm.scratch[warp];        // 13 x 12
      double* Blk = sm.scratch[warp] + 160;  // 12 x 12
      for (int blk = warp; blk < H * (H + 1) / 2; blk += kGenBuildThreads / 32) {
        int j = 0, rem = blk;
        while (rem > j) { rem -= (j + 1); ++j; }
        const int l = rem;
        const double* Ap = &sm.Apow[(j - l) * 169];
        const double* Bl = &sm.Bd[l * 156];
        const double* Tj = &sm.T[j * 156];
        __syncwarp();
        for (int e = lane; e < 156; e += 32) {
          const int rr = e / 12, cc = e - 12 * rr;
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < 13; ++k) a = fma(Ap[rr * 13 + k], Bl[k * 12 + cc], a);
          U[e] = a;
        }
        __syncwarp();
        for (int e = lane; e < 144; e += 32) {
          int ia = e / 12, ib = e - 12 * ia;
          if (j == l && ib > ia) { const int t = ia; ia = ib; ib = t; }  // diagonal block: exact symmetry
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < 13; ++k) a = fma(Tj[ia * 13 + k], U[k * 12 + ib], a);
          if (j == l && ia == ib) a += bp.Rd[ia];
          Blk[e] = a;
        }
        __syncwarp();
        if (lane < 12) {
          double2* dst = reinterpret_cast<double2*>(Pp + size_t(12 * j + lane) * n + 12 * l);
#pragma unroll
          for (int h = 0; h < 6; ++h) dst[h] = make_double2(Blk[lane * 12 + 2 * h], Blk[lane * 12 + 2 * h + 1]);
        } else if (lane >= 16 && lane < 28 && j != l) {
          const int ib = lane - 16;  // the mirrored block: row ib of block (l, j) is column ib of Blk
          double2* dst = reinterpret_cast<double2*>(Pp + size_t(12 * l + ib) * n + 12 * j);
#pragma unroll
          for (int h = 0; h < 6; ++h) dst[h] = make_double2(Blk[(2 * h) * 12 + ib], Blk[(2 * h + 1) * 12 + ib]);
        }
      }
    }
    // bounds
    for (int i = tid; i < m; i += kGenBuildThreads) {
      const int leg = (i % 20) / 5, t = i % 5;
      const float cflag = sm.contacts[4 * (i / 20) + leg] ? 1.0f : 0.0f;
      float lo, hi;
      if (t == 0 || t == 2) { lo = 0.0f; hi = (float)MPC_INFTY; }
      else if (t == 1 || t == 3) { lo = -(float)MPC_INFTY; hi = 0.0f; }
      else { lo = (float)bp.fz_min * cflag; hi = (float)bp.fz_max * cflag; }
      l_out[size_t(p) * m + i] = lo;
      u_out[size_t(p) * m + i] = hi;
    }
  }
}

// ---------------------------------------------------------------------------
template <int H>
struct GenSolveSmem {
  static constexpr int n = 12 * H, m = 20 * H, nls = 4 * H;
  double x[n], xt[n], rhs[n], qb[n], D[n], Dinv[n], q0[n], xD[n], Px[n];
  double z[m], y[m], lb[m], ub[m], E[m], Einv[m], rv[m], rinv[m], w[m];
  double Av[nls * 9];
  double G[nls * 9];
  double V[3][n], W[3][n];
  double Minv[9];
  double red[kGenSolveWarps * 16];
  double scal[8];  // 0:c 1:cinv 2:rho 3:ct 4:pri_res
  int flags[8];    // 0:done 1:status 2:refactor 3:problem index
  int ctype[m];
};

__device__ __forceinline__ double gen_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double gen_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double gen_limit(double v) {
  v = v < 1e-4 ? 1.0 : v;
  return v > 1e4 ? 1e4 : v;
}

// Build K into the workspace and overwrite it with -K^-1 by the blocked symmetric sweep
// (same formulation as admm_kernel.cuh, matrix streamed from the L2-resident workspace).
template <int H>
__device__ void gen_factor(GenSolveSmem<H>& sm, const double* __restrict__ P, double* __restrict__ K, int tid,
                           double sigma) {
  constexpr int n = 12 * H, nls = 4 * H;
  const int lane = tid & 31, warp = tid >> 5;
  // G = A' diag(rho) A, 3x3 per leg-step
  for (int idx = tid; idx < nls * 9; idx += kGenSolveThreads) {
    const int k = idx / 9, rr = (idx % 9) / 3, cc = idx % 3;
    const double* av = &sm.Av[k * 9];
    double g = 0.0;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      double co[3];
      co[0] = (e == 0) ? av[0] : (e == 1) ? av[2] : 0.0;
      co[1] = (e == 2) ? av[4] : (e == 3) ? av[6] : 0.0;
      co[2] = (e == 0) ? av[1] : (e == 1) ? av[3] : (e == 2) ? av[5] : (e == 3) ? av[7] : av[8];
      g += sm.rv[5 * k + e] * co[rr] * co[cc];
    }
    sm.G[idx] = g;
  }
  __syncthreads();
  const double c = sm.scal[0];
  for (int r = warp; r < n; r += kGenSolveWarps) {
    const double cDr = c * sm.D[r];
    for (int j = lane; j < n; j += 32) {
      double v = cDr * P[size_t(r) * n + j] * sm.D[j];
      if (j == r) v += sigma;
      if (j / 3 == r / 3) v += sm.G[(r / 3) * 9 + (r % 3) * 3 + (j % 3)];
      K[size_t(r) * n + j] = v;
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int kb = 0; kb < nls; ++kb) {
    const int c0 = 3 * kb;
    // pivot rows -> smem
    for (int idx = tid; idx < 3 * n; idx += kGenSolveThreads) sm.V[idx / n][idx % n] = K[size_t(c0 + idx / n) * n + idx % n];
    __syncthreads();
    if (tid == 0) {
      const double m00 = sm.V[0][c0], m01 = sm.V[0][c0 + 1], m02 = sm.V[0][c0 + 2];
      const double m11 = sm.V[1][c0 + 1], m12 = sm.V[1][c0 + 2], m22 = sm.V[2][c0 + 2];
      const double k00 = m11 * m22 - m12 * m12, k01 = m02 * m12 - m01 * m22, k02 = m01 * m12 - m02 * m11;
      const double id = 1.0 / (m00 * k00 + m01 * k01 + m02 * k02);
      sm.Minv[0] = k00 * id; sm.Minv[1] = k01 * id; sm.Minv[2] = k02 * id;
      sm.Minv[3] = k01 * id; sm.Minv[4] = (m00 * m22 - m02 * m02) * id; sm.Minv[5] = (m01 * m02 - m00 * m12) * id;
      sm.Minv[6] = k02 * id; sm.Minv[7] = sm.Minv[5]; sm.Minv[8] = (m00 * m11 - m01 * m01) * id;
    }
    __syncthreads();
    for (int j = tid; j < n; j += kGenSolveThreads) {
      const double x0 = sm.V[0][j], x1 = sm.V[1][j], x2 = sm.V[2][j];
#pragma unroll
      for (int s3 = 0; s3 < 3; ++s3)
        sm.W[s3][j] = -(sm.Minv[3 * s3] * x0 + sm.Minv[3 * s3 + 1] * x1 + sm.Minv[3 * s3 + 2] * x2);
    }
    __syncthreads();
    if (tid < 3) sm.V[tid][c0 + tid] -= 1.0;  // V' = V with A_SS - I
    __syncthreads();
    for (int r = warp; r < n; r += kGenSolveWarps) {
      double* Kr = K + size_t(r) * n;
      if (r >= c0 && r < c0 + 3) {
        const int s3 = r - c0;
        for (int j = lane; j < n; j += 32) {
          const int t = j - c0;
          Kr[j] = (t >= 0 && t < 3) ? -sm.Minv[3 * s3 + t] : -sm.W[s3][j];
        }
      } else {
        const double w0 = sm.W[0][r], w1 = sm.W[1][r], w2 = sm.W[2][r];
        for (int j = lane; j < n; j += 32)
          Kr[j] = fma(w0, sm.V[0][j], fma(w1, sm.V[1][j], fma(w2, sm.V[2][j], Kr[j])));
      }
    }
    __threadfence_block();
    __syncthreads();
  }
}

template <int H>
__global__ void __launch_bounds__(kGenSolveThreads, 1)
gen_solve_kernel(const double* __restrict__ P_all, const double* __restrict__ q_all,
                 const float* __restrict__ l_all, const float* __restrict__ u_all,
                 const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                 float* __restrict__ x_all, int num, int* __restrict__ counter, double* __restrict__ workspace,
                 const __grid_constant__ SolveParams sp) {
  constexpr int n = 12 * H, m = 20 * H, nls = 4 * H;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GenSolveSmem<H>& sm = *reinterpret_cast<GenSolveSmem<H>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* K = workspace + size_t(blockIdx.x) * n * n;  // this CTA's -K^-1
  const double mu = sp.mu, sigma = sp.sigma, alpha = sp.alpha;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;
    const double* P = P_all + size_t(p) * n * n;

    for (int j = tid; j < n; j += kGenSolveThreads) {
      sm.q0[j] = q_all[size_t(p) * n + j];
      sm.D[j] = 1.0;
      sm.x[j] = 0.0;
      sm.xt[j] = 0.0;
    }
    for (int i = tid; i < m; i += kGenSolveThreads) {
      sm.lb[i] = (double)l_all[size_t(p) * m + i];
      sm.ub[i] = (double)u_all[size_t(p) * m + i];
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
    __syncthreads();

    // ---- Ruiz equilibration (osqp scaling.c scale_data), scaled matrices never materialised ----
    // row norm pass: Px[r] <- max_j |P_rj| D_j  (column norms of the symmetric P)
    auto norm_pass = [&]() {
      for (int r = warp; r < n; r += kGenSolveWarps) {
        double mm = 0.0;
        for (int j = lane; j < n; j += 32) mm = fmax(mm, fabs(P[size_t(r) * n + j]) * sm.D[j]);
        mm = gen_warp_max(mm);
        if (lane == 0) sm.Px[r] = mm;
      }
    };
    if (sp.scaling > 0) {
      norm_pass();
      __syncthreads();
      // xD holds the current column norm of P_bar
      for (int j = tid; j < n; j += kGenSolveThreads) sm.xD[j] = sm.Px[j];
      __syncthreads();
      for (int it = 0; it < sp.scaling; ++it) {
        // new scalings from the OLD D, E into rhs (D) and w (E)
        for (int j = tid; j < n; j += kGenSolveThreads) {
          const int k = j / 3, c3 = j % 3;
          const double* Ek = &sm.E[5 * k];
          double nA;
          if (c3 == 0) nA = fmax(Ek[0], Ek[1]);
          else if (c3 == 1) nA = fmax(Ek[2], Ek[3]);
          else nA = fmax(mu * fmax(fmax(Ek[0], Ek[1]), fmax(Ek[2], Ek[3])), Ek[4]);
          nA *= sm.D[j];
          sm.rhs[j] = sm.D[j] * rsqrt(gen_limit(fmax(sm.xD[j], nA)));
        }
        for (int i = tid; i < m; i += kGenSolveThreads) {
          const int k = i / 5, e = i % 5;
          const double dz = sm.D[3 * k + 2];
          const double nrow = (e == 4) ? dz : fmax(sm.D[3 * k + ((e < 2) ? 0 : 1)], mu * dz);
          sm.w[i] = sm.E[i] * rsqrt(gen_limit(sm.E[i] * nrow));
        }
        __syncthreads();
        for (int j = tid; j < n; j += kGenSolveThreads) sm.D[j] = sm.rhs[j];
        for (int i = tid; i < m; i += kGenSolveThreads) sm.E[i] = sm.w[i];
        __syncthreads();
        norm_pass();
        __syncthreads();
        const double c_old = sm.scal[0];
        double ps = 0.0, pq = 0.0;
        for (int j = tid; j < n; j += kGenSolveThreads) {
          const double nP2 = c_old * sm.D[j] * sm.Px[j];
          sm.xD[j] = nP2;
          ps += nP2;
          pq = fmax(pq, fabs(c_old * sm.D[j] * sm.q0[j]));
        }
        ps = gen_warp_sum(ps);
        pq = gen_warp_max(pq);
        if (lane == 0) {
          sm.red[warp * 16 + 0] = ps;
          sm.red[warp * 16 + 1] = pq;
        }
        __syncthreads();
        if (tid == 0) {
          double ssum = 0.0, qn = 0.0;
          for (int w = 0; w < kGenSolveWarps; ++w) {
            ssum += sm.red[w * 16 + 0];
            qn = fmax(qn, sm.red[w * 16 + 1]);
          }
          const double ct = 1.0 / gen_limit(fmax(ssum / (double)n, gen_limit(qn)));
          sm.scal[3] = ct;
          sm.scal[0] = c_old * ct;
        }
        __syncthreads();
        for (int j = tid; j < n; j += kGenSolveThreads) sm.xD[j] *= sm.scal[3];
        __syncthreads();
      }
    }
    // ---- scaled data ----
    {
      const double c = sm.scal[0];
      if (tid == 0) sm.scal[1] = 1.0 / c;
      for (int j = tid; j < n; j += kGenSolveThreads) {
        sm.qb[j] = c * sm.D[j] * sm.q0[j];
        sm.Dinv[j] = 1.0 / sm.D[j];
      }
      for (int i = tid; i < m; i += kGenSolveThreads) {
        const double e = sm.E[i];
        const double l = e * sm.lb[i], u = e * sm.ub[i];
        sm.lb[i] = l;
        sm.ub[i] = u;
        sm.Einv[i] = 1.0 / e;
        int ct = 0;
        if (l < -MPC_INFTY * 1e-4 && u > MPC_INFTY * 1e-4) ct = -1;
        else if (u - l < 1e-4) ct = 1;
        sm.ctype[i] = ct;
        const double rvv = (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * sp.rho : sp.rho;
        sm.rv[i] = rvv;
        sm.rinv[i] = 1.0 / rvv;
      }
      for (int k = tid; k < nls; k += kGenSolveThreads) {
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
    for (int j = tid; j < n; j += kGenSolveThreads) sm.rhs[j] = -sm.qb[j];
    gen_factor<H>(sm, P, K, tid, sigma);

    // rhs_j = sigma x_j - q_j + (A'(rho z - y))_j from w = rho z - y
    auto build_rhs = [&]() {
      for (int j = tid; j < n; j += kGenSolveThreads) {
        const int k = j / 3, c3 = j % 3;
        const double* av = &sm.Av[9 * k];
        const double* w = &sm.w[5 * k];
        double sacc;
        if (c3 == 0) sacc = av[0] * w[0] + av[2] * w[1];
        else if (c3 == 1) sacc = av[4] * w[2] + av[6] * w[3];
        else sacc = av[1] * w[0] + av[3] * w[1] + av[5] * w[2] + av[7] * w[3] + av[8] * w[4];
        sm.rhs[j] = sigma * sm.x[j] - sm.qb[j] + sacc;
      }
    };

    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    for (iter = 1; iter <= sp.max_iter; ++iter) {
      // x~ = K^-1 rhs (K holds -K^-1), warp per row over the L2-resident workspace
      for (int r = warp; r < n; r += kGenSolveWarps) {
        const double* Kr = K + size_t(r) * n;
        double a0 = 0.0, a1 = 0.0;
        int j = lane;
        for (; j + 32 < n; j += 64) {
          a0 = fma(Kr[j], sm.rhs[j], a0);
          a1 = fma(Kr[j + 32], sm.rhs[j + 32], a1);
        }
        if (j < n) a0 = fma(Kr[j], sm.rhs[j], a0);
        const double tot = gen_warp_sum(a0 + a1);
        if (lane == 0) {
          const double xt = -tot;
          sm.xt[r] = xt;
          sm.x[r] = alpha * xt + (1.0 - alpha) * sm.x[r];
        }
      }
      __syncthreads();
      for (int i = tid; i < m; i += kGenSolveThreads) {
        const int k = i / 5, e = i % 5;
        const double* av = &sm.Av[9 * k];
        const double cca = (e < 4) ? av[2 * e] : 0.0, ccz = (e < 4) ? av[2 * e + 1] : av[8];
        const double zt = cca * sm.xt[3 * k + ((e < 2) ? 0 : 1)] + ccz * sm.xt[3 * k + 2];
        const double zr = alpha * zt + (1.0 - alpha) * sm.z[i];
        const double rvv = sm.rv[i], yo = sm.y[i];
        double zn = zr + sm.rinv[i] * yo;
        zn = fmin(fmax(zn, sm.lb[i]), sm.ub[i]);
        const double yn = yo + rvv * (zr - zn);
        sm.z[i] = zn;
        sm.y[i] = yn;
        sm.w[i] = rvv * zn - yn;
      }
      __syncthreads();
      build_rhs();
      const bool can_check = (--until_check == 0);
      const bool can_adapt = (--until_adapt == 0);
      if (can_check) until_check = sp.check_termination;
      if (can_adapt) until_adapt = sp.adaptive_rho_interval;
      const bool last = (iter == sp.max_iter);
      if (!(can_check || can_adapt || last)) {
        __syncthreads();
        continue;
      }
      // ---- residuals ----
      for (int j = tid; j < n; j += kGenSolveThreads) sm.xD[j] = sm.D[j] * sm.x[j];
      __syncthreads();
      for (int r = warp; r < n; r += kGenSolveWarps) {
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) acc = fma(P[size_t(r) * n + j], sm.xD[j], acc);
        acc = gen_warp_sum(acc);
        if (lane == 0) sm.Px[r] = sm.scal[0] * sm.D[r] * acc;
      }
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      for (int i = tid; i < m; i += kGenSolveThreads) {
        const int k = i / 5, e = i % 5;
        const double* av = &sm.Av[9 * k];
        const double cca = (e < 4) ? av[2 * e] : 0.0, ccz = (e < 4) ? av[2 * e + 1] : av[8];
        const double Ax = cca * sm.x[3 * k + ((e < 2) ? 0 : 1)] + ccz * sm.x[3 * k + 2];
        const double zz = sm.z[i], rp_ = Ax - zz, ei = sm.Einv[i];
        v[0] = fmax(v[0], fabs(rp_));
        v[1] = fmax(v[1], fabs(ei * rp_));
        v[2] = fmax(v[2], fabs(ei * zz));
        v[3] = fmax(v[3], fabs(ei * Ax));
        v[4] = fmax(v[4], fabs(zz));
        v[5] = fmax(v[5], fabs(Ax));
      }
      for (int j = tid; j < n; j += kGenSolveThreads) {
        const int k = j / 3, c3 = j % 3;
        const double* av = &sm.Av[9 * k];
        const double* yy = &sm.y[5 * k];
        double Aty;
        if (c3 == 0) Aty = av[0] * yy[0] + av[2] * yy[1];
        else if (c3 == 1) Aty = av[4] * yy[2] + av[6] * yy[3];
        else Aty = av[1] * yy[0] + av[3] * yy[1] + av[5] * yy[2] + av[7] * yy[3] + av[8] * yy[4];
        const double Px = sm.Px[j], qq = sm.qb[j], rd = Px + qq + Aty, di = sm.Dinv[j];
        v[6] = fmax(v[6], fabs(rd));
        v[7] = fmax(v[7], fabs(di * rd));
        v[8] = fmax(v[8], fmax(fmax(fabs(di * qq), fabs(di * Aty)), fabs(di * Px)));
        v[9] = fmax(v[9], fmax(fmax(fabs(qq), fabs(Aty)), fabs(Px)));
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const double mm = gen_warp_max(v[i]);
        if (lane == 0) sm.red[warp * 16 + i] = mm;
      }
      __syncthreads();
      if (tid == 0) {
        double mx[10];
        for (int i = 0; i < 10; ++i) {
          double t = 0.0;
          for (int w = 0; w < kGenSolveWarps; ++w) t = fmax(t, sm.red[w * 16 + i]);
          mx[i] = t;
        }
        const double cinv = sm.scal[1];
        const double pri = mx[1], dua = cinv * mx[7];
        const double eps_pri = sp.eps_abs + sp.eps_rel * fmax(mx[2], mx[3]);
        const double eps_dua = sp.eps_abs + sp.eps_rel * cinv * mx[8];
        sm.scal[4] = pri;
        int done = 0, refactor = 0;
        if ((can_check || last) && pri < eps_pri && dua < eps_dua) {
          done = 1;
          sm.flags[1] = MPC_STATUS_SOLVED;
        } else if (last) {
          done = 1;
          sm.flags[1] = (pri < 10.0 * eps_pri && dua < 10.0 * eps_dua) ? 2 : MPC_STATUS_MAX_ITER_REACHED;
        } else if (can_adapt) {
          const double rho = sm.scal[2];
          const double pn = mx[0] / (fmax(mx[4], mx[5]) + 1e-10);
          const double dn = mx[6] / (mx[9] + 1e-10);
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
        for (int i = tid; i < m; i += kGenSolveThreads) {
          const int ct = sm.ctype[i];
          const double rvv = (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho;
          sm.rv[i] = rvv;
          sm.rinv[i] = 1.0 / rvv;
          sm.w[i] = rvv * sm.z[i] - sm.y[i];
        }
        __syncthreads();
        build_rhs();
        __syncthreads();
        gen_factor<H>(sm, P, K, tid, sigma);
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    __syncthreads();
    if (x_all != nullptr)
      for (int j = tid; j < n; j += kGenSolveThreads) x_all[size_t(p) * n + j] = (float)(sm.D[j] * sm.x[j]);
    if (tid < 12) {
      const int leg = tid / 3, rr = tid % 3;
      const double f0 = sm.D[3 * leg] * sm.x[3 * leg], f1 = sm.D[3 * leg + 1] * sm.x[3 * leg + 1],
                   f2 = sm.D[3 * leg + 2] * sm.x[3 * leg + 2];
      double g;
      if (states != nullptr) {
        const float* R = reinterpret_cast<const float*>(states + p) + kOffRot;
        g = (double)R[rr] * f0 + (double)R[3 + rr] * f1 + (double)R[6 + rr] * f2;
      } else {
        g = (rr == 0) ? f0 : (rr == 1) ? f1 : f2;
      }
      const bool bad = isnan(f0) || isnan(f1) || isnan(f2);
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
