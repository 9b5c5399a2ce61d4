// K3+K4+K5: OSQP-equivalent ADMM solve of the condensed MPC QP (H = 10), one CTA per
// problem, problems pulled from an atomic counter.  Restates OSQP 0.6.x as the reference
// drives it (A1RobotControl.cpp:522-561): modified Ruiz equilibration, per-row rho,
// K = P + sigma I + A' diag(rho) A, ADMM with alpha-relaxation, residual termination
// every check_termination iterations, rho adaptation with refactorisation.
//
// v3 layout (v1/v2 history in profiles/):
//   640 threads = 40 row groups x 16 column groups; thread (rg, cg) holds the 3 x 8
//   register tile rows 3rg..3rg+2 x columns {32i + 2cg, 32i + 2cg + 1 : i < 4} of -K^-1.
//   A row group IS a leg-step: variables 3rg..3rg+2, constraint rows 5rg..5rg+4.  The 16
//   lanes of a half-warp therefore finish x~ for one leg-step (reduce-scatter), and the
//   same lanes do that leg-step's z/y update and next right-hand side through shuffles:
//   ONE block barrier per ADMM iteration, no shared-memory vectors except rhs.
//     lane cg = 0 : variable fx + row 0      lane cg = 1 : row 1
//     lane cg = 4 : variable fy + row 2      lane cg = 5 : row 3
//     lane cg = 8 : variable fz + row 4
//   K^-1 comes from a BLOCKED symmetric sweep (Gauss-Jordan on the SPD matrix) over the
//   register tiles: a row group is exactly 3 pivot rows, so one rank-3 update per barrier;
//   the publisher stores only its three rows, every thread inverts the 3x3 pivot block itself.
//   (One pivot per barrier, or M/W computed by the publisher alone, left 40-50 % of the sweep
//   time at the barrier: profiles/r01_v3_*, r01_v4_*.)
//   P arrives by one cp.async.bulk (TMA) per problem into shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_kernels.cuh"

namespace mpcb200 {

constexpr int kSolveThreads = 640;
constexpr int kSolveWarps = kSolveThreads / 32;
constexpr int kPBytes = kN * kNP * 8;  // 122,880

struct SolveSmem {
  double P[kN * kNP];       // unscaled Hessian, row stride 128, pad columns zero (TMA destination)
  double rhs[2][kNP];       // operand of the K^-1 matvec, double buffered by iteration parity (pad = 0)
  double xD[kNP];           // D .* x for P x (pad = 0)
  double Dp[kNP];           // D (pad = 0)
  double Vb[2][3][kNP];     // blocked sweep: published pivot rows A_S,: with A_SS replaced by A_SS - I
  double G[kLegSteps * 6];  // A' diag(rho) A per leg-step: xx, xz, yy, yz, zz, (pad)
  // per-lane constants of the ADMM loop (slot [k][tid]); registers are kept for the K^-1 tile
  double lane_lb[kSolveThreads], lane_ub[kSolveThreads], lane_rv[kSolveThreads], lane_rinv[kSolveThreads];
  double lane_qb[kSolveThreads], lane_D[kSolveThreads], lane_Einv[kSolveThreads];
  double red[kSolveWarps * 16];
  double scal[8];           // 0:c 1:cinv 2:rho 3:ct 4:pri_res
  unsigned long long mbar;
  int flags[8];             // 0:done 1:status 2:refactor 3:problem index
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
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Reduce three per-row partials over the 16 lanes of a half-warp: lanes 4c..4c+3 end up
// with the total of row c (c = 0, 1, 2); lanes 12..15 hold nothing useful.
__device__ __forceinline__ double reduce_scatter3_sum(double s0, double s1, double s2, int cg) {
  const bool h8 = (cg & 8) != 0, h4 = (cg & 4) != 0;
  double k0 = h8 ? s2 : s0, k1 = h8 ? 0.0 : s1;
  const double t0 = h8 ? s0 : s2, t1 = h8 ? s1 : 0.0;
  k0 += __shfl_xor_sync(0xffffffffu, t0, 8);
  k1 += __shfl_xor_sync(0xffffffffu, t1, 8);
  double k = h4 ? k1 : k0;
  const double t = h4 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, t, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
__device__ __forceinline__ double reduce_scatter3_max(double s0, double s1, double s2, int cg) {
  const bool h8 = (cg & 8) != 0, h4 = (cg & 4) != 0;
  double k0 = h8 ? s2 : s0, k1 = h8 ? 0.0 : s1;
  const double t0 = h8 ? s0 : s2, t1 = h8 ? s1 : 0.0;
  k0 = fmax(k0, __shfl_xor_sync(0xffffffffu, t0, 8));
  k1 = fmax(k1, __shfl_xor_sync(0xffffffffu, t1, 8));
  double k = h4 ? k1 : k0;
  const double t = h4 ? k0 : k1;
  k = fmax(k, __shfl_xor_sync(0xffffffffu, t, 4));
  k = fmax(k, __shfl_xor_sync(0xffffffffu, k, 2));
  k = fmax(k, __shfl_xor_sync(0xffffffffu, k, 1));
  return k;
}

// Sums over the five row lanes {0,1,4,5,8} of a half-warp (hb = lane & 16):
//   returns on lane 0: pa@0 + pa@1 (in .x) ; lane 4: pa@4 + pa@5 (in .x) ; lane 8: sum of pz over all five (in .y)
struct LegSums { double lat, z; };
__device__ __forceinline__ LegSums leg_reduce(double pa, double pz, int hb) {
  LegSums r;
  r.lat = pa + __shfl_down_sync(0xffffffffu, pa, 1);
  const double t = pz + __shfl_down_sync(0xffffffffu, pz, 1);
  r.z = pz + shfl(t, hb) + shfl(t, hb + 4);
  return r;
}

__device__ __forceinline__ void load_cols(const double* base, int cg, double (&v)[8]) {
  const double2* p = reinterpret_cast<const double2*>(base + 2 * cg);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 t = p[16 * i];
    v[2 * i] = t.x;
    v[2 * i + 1] = t.y;
  }
}

// max_j |P_rj| D_j of the rows of this row group; lanes 4c..4c+3 get row c
__device__ __forceinline__ double row_norm_pass(const SolveSmem& sm, int rg, int cg) {
  double dcol[8];
  load_cols(sm.Dp, cg, dcol);
  double m[3];
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) {
    double pv[8];
    load_cols(&sm.P[(3 * rg + rr) * kNP], cg, pv);
    double mm = 0.0;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) mm = fmax(mm, fabs(pv[jj]) * dcol[jj]);
    m[rr] = mm;
  }
  return reduce_scatter3_max(m[0], m[1], m[2], cg);
}

// Blocked symmetric sweep, one leg-step (3 pivots S = {3kb, 3kb+1, 3kb+2}) per barrier.
// With V = A_S,: (before the step), M = A_SS^-1, W = -M V the step is
//   A_rj <- A_rj + sum_s W[s][r] V'[s][j]   (r not in S; V' = V with A_SS - I in the S columns,
//                                            which makes the same update produce A_rS M)
//   A_Sj <- -W[:, j] (j not in S),  A_SS <- -M
// (W[s][r] doubles as the column factor because A is symmetric).  The row group kb IS S, so
// its 16 lanes hold V entirely and publish ONLY V' (12 stores): every thread inverts the 3x3
// pivot block itself and forms the 3x3 block of W it needs.  Doing M and W on the publishing
// half-warp alone left the other 19 warps at the barrier for half of every step.
__device__ __forceinline__ void publish_rows(SolveSmem& sm, const double (&a)[3][8], int kb, int cg, int b) {
  const int c0 = 3 * kb;
#pragma unroll
  for (int s3 = 0; s3 < 3; ++s3) {
    double2* dst = reinterpret_cast<double2*>(&sm.Vb[b][s3][2 * cg]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = 32 * i + 2 * cg;
      double v0 = a[s3][2 * i], v1 = a[s3][2 * i + 1];
      if (col == c0 + s3) v0 -= 1.0;
      if (col + 1 == c0 + s3) v1 -= 1.0;
      dst[16 * i] = make_double2(v0, v1);
    }
  }
}

// One step of the blocked sweep for this thread's tile, block kb published in buffer b.
__device__ __forceinline__ void sweep_block(const SolveSmem& sm, double (&a)[3][8], int kb, int rg, int cg,
                                            int b) {
  const int c0 = 3 * kb;
  // A_SS (identity added back) and its symmetric 3x3 cofactor inverse, on every thread
  const double m00 = sm.Vb[b][0][c0] + 1.0, m01 = sm.Vb[b][0][c0 + 1], m02 = sm.Vb[b][0][c0 + 2];
  const double m11 = sm.Vb[b][1][c0 + 1] + 1.0, m12 = sm.Vb[b][1][c0 + 2];
  const double m22 = sm.Vb[b][2][c0 + 2] + 1.0;
  const double k00 = m11 * m22 - m12 * m12, k01 = m02 * m12 - m01 * m22, k02 = m01 * m12 - m02 * m11;
  const double id = __drcp_rn(m00 * k00 + m01 * k01 + m02 * k02);
  const double i00 = k00 * id, i01 = k01 * id, i02 = k02 * id;
  const double i11 = (m00 * m22 - m02 * m02) * id, i12 = (m01 * m02 - m00 * m12) * id;
  const double i22 = (m00 * m11 - m01 * m01) * id;
  if (rg == kb) {
    // pivot rows: A_Sj <- M A_Sj, A_SS <- -M
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double x0 = a[0][jj], x1 = a[1][jj], x2 = a[2][jj];
      const int t = 32 * (jj >> 1) + 2 * cg + (jj & 1) - c0;  // position inside S, if any
      a[0][jj] = (t == 0) ? -i00 : (t == 1) ? -i01 : (t == 2) ? -i02 : (i00 * x0 + i01 * x1 + i02 * x2);
      a[1][jj] = (t == 0) ? -i01 : (t == 1) ? -i11 : (t == 2) ? -i12 : (i01 * x0 + i11 * x1 + i12 * x2);
      a[2][jj] = (t == 0) ? -i02 : (t == 1) ? -i12 : (t == 2) ? -i22 : (i02 * x0 + i12 * x1 + i22 * x2);
    }
  } else {
    // W[s][r] = -(M V[:, r])[s] for the thread's three rows (V = V' there: r is not in S)
    double w[3][3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      const double x0 = sm.Vb[b][0][3 * rg + rr], x1 = sm.Vb[b][1][3 * rg + rr], x2 = sm.Vb[b][2][3 * rg + rr];
      w[0][rr] = -(i00 * x0 + i01 * x1 + i02 * x2);
      w[1][rr] = -(i01 * x0 + i11 * x1 + i12 * x2);
      w[2][rr] = -(i02 * x0 + i12 * x1 + i22 * x2);
    }
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3) {
      double v[8];
      load_cols(sm.Vb[b][s3], cg, v);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        a[0][jj] = fma(w[s3][0], v[jj], a[0][jj]);
        a[1][jj] = fma(w[s3][1], v[jj], a[1][jj]);
        a[2][jj] = fma(w[s3][2], v[jj], a[2][jj]);
      }
    }
  }
}

// Build K = c D P D + sigma I + A' diag(rho) A into the register tiles, then overwrite it
// with -K^-1 by the blocked symmetric sweep (40 rank-3 steps, one barrier each).
__device__ __forceinline__ void factor_inverse(SolveSmem& sm, double (&a)[3][8], int rg, int cg, int hb,
                                               double sigma) {
  {
    const double c = sm.scal[0];
    double dcol[8];
    load_cols(sm.Dp, cg, dcol);
    const double* g = &sm.G[rg * 6];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      const int row = 3 * rg + rr;
      const double cDr = c * sm.Dp[row];
      double pv[8];
      load_cols(&sm.P[row * kNP], cg, pv);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int col = 32 * (jj >> 1) + 2 * cg + (jj & 1);
        double val = cDr * pv[jj] * dcol[jj];
        if (col == row) val += sigma;
        if (col / 3 == rg) {
          const int cc = col - 3 * rg;
          // G = [[xx, 0, xz], [0, yy, yz], [xz, yz, zz]]
          double gv;
          if (rr == 0) gv = (cc == 0) ? g[0] : (cc == 1) ? 0.0 : g[1];
          else if (rr == 1) gv = (cc == 0) ? 0.0 : (cc == 1) ? g[2] : g[3];
          else gv = (cc == 0) ? g[1] : (cc == 1) ? g[3] : g[4];
          val += gv;
        }
        a[rr][jj] = val;
      }
    }
  }
  if (rg == 0) publish_rows(sm, a, 0, cg, 0);
  for (int kb = 0; kb < kLegSteps; ++kb) {
    const int b = kb & 1;
    __syncthreads();  // rows of block kb (published one step ahead) are visible
    sweep_block(sm, a, kb, rg, cg, b);
    // look-ahead: the next pivot rows are now up to date; publish them for step kb + 1
    if (rg == kb + 1) publish_rows(sm, a, kb + 1, cg, b ^ 1);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <bool kProfile>
__global__ void __launch_bounds__(kSolveThreads, 1)
admm_solve_kernel(const double* __restrict__ P_all, const double* __restrict__ q_all,
                  const float* __restrict__ l_all, const float* __restrict__ u_all,
                  const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                  float* __restrict__ x_all, int num, int* __restrict__ counter,
                  long long* __restrict__ phase_clk, const __grid_constant__ SolveParams sp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SolveSmem& sm = *reinterpret_cast<SolveSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = tid >> 4, cg = tid & 15, hb = lane & 16;
  // lane roles inside the half-warp (= leg-step rg)
  const bool vown = (cg == 0) || (cg == 4) || (cg == 8);
  const int vc = cg >> 2;                 // variable component owned (x, y, z)
  const int vj = 3 * rg + (vown ? vc : 0);
  const bool rown = (cg == 0) || (cg == 1) || (cg == 4) || (cg == 5) || (cg == 8);
  const int re = (cg & 1) + ((cg >> 2) << 1);  // row inside the leg-step (valid on row lanes)
  const int ri = 5 * rg + (rown ? re : 0);
  const int latsrc = hb + (cg & 4), zsrc = hb + 8;
  const double mu = sp.mu;
  const double sigma = sp.sigma, alpha = sp.alpha;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;
  double a[3][8];  // register tile of -K^-1
  // per-phase cycle counters of thread 0: compiled in only for the kProfile instantiation
  // (they cost 14 registers the production kernel cannot spare)
  long long pc[kProfile ? 6 : 1] = {0};  // 0 load+ruiz 1 factor 2 iterations 3 checks 4 output 5 problems
  long long tmark = 0;
#define PHASE_MARK(i) do { if (kProfile && tid == 0) { const long long _t = clock64(); pc[kProfile ? (i) : 0] += _t - tmark; tmark = _t; } } while (0)

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const int pnext = atomicAdd(counter, 1);
      sm.flags[3] = pnext;
      if (pnext < num) {
        // K3 loader: one TMA bulk copy brings the whole padded Hessian into shared memory
        const uint32_t bar = smem_u32(&sm.mbar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPBytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(sm.P)),
            "l"(P_all + size_t(pnext) * kN * kNP), "r"(kPBytes), "r"(bar)
            : "memory");
      }
    }
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;
    if (kProfile && tid == 0) { tmark = clock64(); pc[kProfile ? 5 : 0] += 1; }

    if (tid < kNP) {
      sm.Dp[tid] = (tid < kN) ? 1.0 : 0.0;
      sm.rhs[0][tid] = 0.0;
      sm.rhs[1][tid] = 0.0;
      sm.xD[tid] = 0.0;
#pragma unroll
      for (int s3 = 0; s3 < 3; ++s3) {
        sm.Vb[0][s3][tid] = 0.0;
        sm.Vb[1][s3][tid] = 0.0;
      }
    }
    if (tid == 0) {
      sm.scal[0] = 1.0;
      sm.scal[2] = sp.rho;
      sm.flags[0] = 0;
      sm.flags[1] = MPC_STATUS_UNSOLVED;
    }
    const double q0 = vown ? q_all[size_t(p) * kN + vj] : 0.0;
    double lb = rown ? (double)l_all[size_t(p) * kM + ri] : 0.0;
    double ub = rown ? (double)u_all[size_t(p) * kM + ri] : 0.0;
    {
      // wait for the bulk copy (phase parity flips once per problem)
      const uint32_t bar = smem_u32(&sm.mbar);
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(phase)
            : "memory");
      }
      phase ^= 1;
    }
    __syncthreads();

    // ---- K3a: modified Ruiz equilibration (osqp scaling.c scale_data) ----
    // scaled quantities are never materialised: P_bar = c D P D, A_bar = E A D.
    double D = 1.0, E = 1.0;  // D on variable lanes, E on row lanes
    if (sp.scaling > 0) {
      double nP = row_norm_pass(sm, rg, cg);  // c = 1, D = 1
      __syncthreads();                        // Dp is rewritten inside the loop
      for (int it = 0; it < sp.scaling; ++it) {
        // column norm of [P; A] for the owned variable
        const double E0 = shfl(E, hb), E1 = shfl(E, hb + 1), E2 = shfl(E, hb + 4), E3 = shfl(E, hb + 5),
                     E4 = shfl(E, hb + 8);
        const double Dx = shfl(D, hb), Dy = shfl(D, hb + 4), Dz = shfl(D, hb + 8);
        if (vown) {
          double nA;
          if (vc == 0) nA = fmax(E0, E1);
          else if (vc == 1) nA = fmax(E2, E3);
          else nA = fmax(mu * fmax(fmax(E0, E1), fmax(E2, E3)), E4);
          nA *= D;
          D *= rsqrt(limit_scaling(fmax(nP, nA)));
          sm.Dp[vj] = D;
        }
        if (rown) {
          // row norm of A for the owned constraint row
          const double nrow = (re == 4) ? Dz : fmax((re < 2) ? Dx : Dy, mu * Dz);
          E *= rsqrt(limit_scaling(E * nrow));
        }
        __syncthreads();
        // cost normalisation with the new D and the old c
        const double c_old = sm.scal[0];
        const double nP2 = c_old * D * row_norm_pass(sm, rg, cg);
        double part_sum = vown ? nP2 : 0.0;
        double part_q = vown ? fabs(c_old * D * q0) : 0.0;
        part_sum = warp_sum(part_sum);
        part_q = warp_max(part_q);
        if (lane == 0) {
          sm.red[warp * 16 + 0] = part_sum;
          sm.red[warp * 16 + 1] = part_q;
        }
        __syncthreads();
        if (tid == 0) {
          double s = 0.0, qn = 0.0;
          for (int w = 0; w < kSolveWarps; ++w) {
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
    // ---- scaled data on the owning lanes (constants go to per-lane shared-memory slots) ----
    const double c = sm.scal[0];
    const double qb0 = c * D * q0;
    lb *= E;
    ub *= E;
    int ctype = 0;
    if (lb < -MPC_INFTY * 1e-4 && ub > MPC_INFTY * 1e-4) ctype = -1;
    else if (ub - lb < 1e-4) ctype = 1;
    double rv = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * sp.rho : sp.rho;
    // scaled constraint coefficients of the owned row: z~_i = cca * x~_lat + ccz * x~_z
    double cca, ccz;
    {
      const double Dlat = shfl(D, latsrc), Dz = shfl(D, zsrc);
      cca = (re == 4) ? 0.0 : E * Dlat;
      ccz = (re == 4) ? E * Dz : ((re & 1) ? -mu : mu) * E * Dz;
      if (!rown) { cca = 0.0; ccz = 0.0; }
    }
    sm.lane_lb[tid] = lb;
    sm.lane_ub[tid] = ub;
    sm.lane_rv[tid] = rv;
    sm.lane_rinv[tid] = 1.0 / rv;
    sm.lane_qb[tid] = qb0;
    sm.lane_D[tid] = D;
    sm.lane_Einv[tid] = 1.0 / E;
    if (tid == 0) sm.scal[1] = 1.0 / c;
    auto build_G = [&]() {
      // A' diag(rho) A of this leg-step from the five row lanes
      const double r0 = rown ? sm.lane_rv[tid] : 0.0;
      const LegSums s1 = leg_reduce(r0 * cca * cca, r0 * ccz * ccz, hb);  // xx|yy on lanes 0|4, zz on lane 8
      const LegSums s2 = leg_reduce(r0 * cca * ccz, 0.0, hb);             // xz|yz on lanes 0|4
      if (cg == 0) { sm.G[rg * 6 + 0] = s1.lat; sm.G[rg * 6 + 1] = s2.lat; }
      if (cg == 4) { sm.G[rg * 6 + 2] = s1.lat; sm.G[rg * 6 + 3] = s2.lat; }
      if (cg == 8) sm.G[rg * 6 + 4] = s1.z;
    };
    build_G();
    if (vown) sm.rhs[1][vj] = -qb0;  // rhs of iteration 1: x = z = y = 0
    __syncthreads();

    PHASE_MARK(0);
    // ---- K3b: factor (explicit inverse in registers) ----
    factor_inverse(sm, a, rg, cg, hb, sigma);
    PHASE_MARK(1);

    // ---- K4: ADMM iterations (osqp.c osqp_solve) ----
    double x = 0.0, z = 0.0, y = 0.0;  // x on variable lanes; z, y on row lanes
    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    // countdowns instead of iter % interval (runtime divisors cost ~50 instructions per iteration)
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    for (iter = 1; iter <= sp.max_iter; ++iter) {
      __syncthreads();  // rhs[iter & 1] is complete; rhs[(iter + 1) & 1] is free to rewrite
      // x~ = K^-1 rhs
      double xt;
      {
        double v[8];
        load_cols(sm.rhs[iter & 1], cg, v);
        double s[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          double s0 = a[rr][0] * v[0], s1 = a[rr][1] * v[1];
#pragma unroll
          for (int jj = 2; jj < 8; jj += 2) {
            s0 = fma(a[rr][jj], v[jj], s0);
            s1 = fma(a[rr][jj + 1], v[jj + 1], s1);
          }
          s[rr] = s0 + s1;
        }
        xt = -reduce_scatter3_sum(s[0], s[1], s[2], cg);  // lanes 4c..4c+3: x~ of variable c
      }
      // x <- alpha x~ + (1 - alpha) x
      x = alpha * xt + (1.0 - alpha) * x;
      // z~ = A x~ ; z, y update on the row lanes
      const double rvv = sm.lane_rv[tid];
      {
        const double xt_lat = shfl(xt, latsrc), xt_z = shfl(xt, zsrc);
        const double zt = cca * xt_lat + ccz * xt_z;
        const double zr = alpha * zt + (1.0 - alpha) * z;
        double zn = zr + sm.lane_rinv[tid] * y;
        zn = fmin(fmax(zn, sm.lane_lb[tid]), sm.lane_ub[tid]);
        y = y + rvv * (zr - zn);
        z = zn;
      }
      // next rhs = sigma x - q + A'(rho z - y)
      {
        const double w = rown ? (rvv * z - y) : 0.0;
        const LegSums s = leg_reduce(cca * w, ccz * w, hb);
        if (vown) sm.rhs[(iter + 1) & 1][vj] = sigma * x - sm.lane_qb[tid] + ((vc == 2) ? s.z : s.lat);
      }
      const bool can_check = (--until_check == 0);
      const bool can_adapt = (--until_adapt == 0);
      if (can_check) until_check = sp.check_termination;
      if (can_adapt) until_adapt = sp.adaptive_rho_interval;
      const bool last = (iter == sp.max_iter);
      if (!(can_check || can_adapt || last)) continue;
      PHASE_MARK(2);

      // ---- residuals (auxil.c compute_pri_res / compute_dua_res / tolerances) ----
      const double Dl = sm.lane_D[tid], cinv = sm.scal[1], c_s = sm.scal[0];
      if (vown) sm.xD[vj] = Dl * x;
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      {
        const double x_lat = shfl(x, latsrc), x_z = shfl(x, zsrc);
        if (rown) {
          const double Einv = sm.lane_Einv[tid];
          const double Ax = cca * x_lat + ccz * x_z;
          const double rp_ = Ax - z;
          v[0] = fabs(rp_);          // scaled primal residual
          v[1] = fabs(Einv * rp_);   // unscaled
          v[2] = fabs(Einv * z);
          v[3] = fabs(Einv * Ax);
          v[4] = fabs(z);
          v[5] = fabs(Ax);
        }
      }
      {
        // P_bar x = c D (P (D x)) ; A' y
        double xv[8];
        load_cols(sm.xD, cg, xv);
        double s[3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          double pv[8];
          load_cols(&sm.P[(3 * rg + rr) * kNP], cg, pv);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int jj = 0; jj < 8; jj += 2) {
            s0 = fma(pv[jj], xv[jj], s0);
            s1 = fma(pv[jj + 1], xv[jj + 1], s1);
          }
          s[rr] = s0 + s1;
        }
        const double sr = reduce_scatter3_sum(s[0], s[1], s[2], cg);
        const double yy = rown ? y : 0.0;
        const LegSums ay = leg_reduce(cca * yy, ccz * yy, hb);
        if (vown) {
          const double qb = sm.lane_qb[tid], Dinv = 1.0 / Dl;
          const double Px = c_s * Dl * sr;
          const double Aty = (vc == 2) ? ay.z : ay.lat;
          const double rd = Px + qb + Aty;
          v[6] = fabs(rd);          // scaled dual residual
          v[7] = fabs(Dinv * rd);   // unscaled (times cinv later)
          v[8] = fmax(fmax(fabs(Dinv * qb), fabs(Dinv * Aty)), fabs(Dinv * Px));
          v[9] = fmax(fmax(fabs(qb), fabs(Aty)), fabs(Px));
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
          for (int w = 0; w < kSolveWarps; ++w) t = fmax(t, sm.red[w * 16 + i]);
          m[i] = t;
        }
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
          const double rho_c = sm.scal[2];
          const double pn = m[0] / (fmax(m[4], m[5]) + 1e-10);
          const double dn = m[6] / (m[9] + 1e-10);
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
      PHASE_MARK(3);
      if (sm.flags[0]) {
        status = sm.flags[1];
        pri_res_out = sm.scal[4];
        break;
      }
      if (sm.flags[2]) {
        ++rho_updates;
        const double rho = sm.scal[2];
        const double rvn = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho : rho;
        sm.lane_rv[tid] = rvn;
        sm.lane_rinv[tid] = 1.0 / rvn;
        // the rhs was built with the old rho vector: rebuild it, then refactor
        {
          const double w = rown ? (rvn * z - y) : 0.0;
          const LegSums s = leg_reduce(cca * w, ccz * w, hb);
          if (vown) sm.rhs[(iter + 1) & 1][vj] = sigma * x - sm.lane_qb[tid] + ((vc == 2) ? s.z : s.lat);
        }
        build_G();
        __syncthreads();
        factor_inverse(sm, a, rg, cg, hb, sigma);
        PHASE_MARK(1);
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    const double xo = sm.lane_D[tid] * x;
    if (x_all != nullptr && vown) x_all[size_t(p) * kN + vj] = (float)xo;
    if (rg < 4) {
      // leg rg of the first horizon step: f = (x, y, z) on lanes 0, 4, 8 of this half-warp
      const double f0 = shfl(xo, hb), f1 = shfl(xo, hb + 4), f2 = shfl(xo, hb + 8);
      if (vown) {
        double g;
        if (states != nullptr) {
          // R' f (A1RobotControl.cpp:558-561)
          const float* R = reinterpret_cast<const float*>(states + p) + kOffRot;
          g = (double)R[vc] * f0 + (double)R[3 + vc] * f1 + (double)R[6 + vc] * f2;
        } else {
          g = (vc == 0) ? f0 : (vc == 1) ? f1 : f2;
        }
        const bool bad = isnan(f0) || isnan(f1) || isnan(f2);  // NaN guard (:559)
        results[p].grf[3 * rg + vc] = bad ? 0.0f : (float)g;
      }
    }
    if (tid == 0) {
      results[p].status = status;
      results[p].iters = iter;
      results[p].rho_updates = rho_updates;
      results[p].pri_res = (float)pri_res_out;
    }
    PHASE_MARK(4);
  }
  if (kProfile && tid == 0 && phase_clk != nullptr) {
#pragma unroll
    for (int i = 0; i < (kProfile ? 6 : 1); ++i) phase_clk[blockIdx.x * 6 + i] = pc[i];
  }
#undef PHASE_MARK
}

}  // namespace mpcb200
