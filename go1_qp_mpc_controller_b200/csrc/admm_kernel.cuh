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
//   K^-1 comes from a symmetric sweep (Gauss-Jordan on the SPD matrix) over the register
//   tiles: per pivot one published row; the row of the NEXT pivot is updated and
//   published first (look-ahead), so the barrier never waits on the publisher.
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
  double buf[2][kNP];       // sweep: published pivot row v' (pivot replaced by d-1)
  double wbuf[2][kNP];      // sweep: -v / d
  double piv[2][2];         // sweep: 1/d
  double G[kLegSteps * 6];  // A' diag(rho) A per leg-step: xx, xz, yy, yz, zz, (pad)
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

// publish row `kn` (the next pivot) held in a[nr][*] by the 16 lanes of its row group
__device__ __forceinline__ void publish_row(SolveSmem& sm, const double (&row)[8], int kn, int cg, int hb,
                                            int nxt) {
  // the pivot d = a[kn][kn] lives on lane (kn >> 1) & 15 at tile position 2 * (kn >> 5) + (kn & 1)
  const int pcg = (kn >> 1) & 15, pi = kn >> 5, pe = kn & 1;
  double cand = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i == pi) cand = pe ? row[2 * i + 1] : row[2 * i];
  // only this half-warp (one row group) executes here: shuffle with its 16-lane mask
  const double d = __shfl_sync(0xffffu << hb, cand, hb + pcg);
  const double dinv = __drcp_rn(d);
  double2* vdst = reinterpret_cast<double2*>(&sm.buf[nxt][2 * cg]);
  double2* wdst = reinterpret_cast<double2*>(&sm.wbuf[nxt][2 * cg]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v0 = row[2 * i], v1 = row[2 * i + 1];
    wdst[16 * i] = make_double2(-v0 * dinv, -v1 * dinv);
    if (cg == pcg && i == pi) {
      if (pe) v1 = d - 1.0; else v0 = d - 1.0;
    }
    vdst[16 * i] = make_double2(v0, v1);
  }
  if (cg == pcg) sm.piv[nxt][0] = dinv;
}

// One pivot of the sweep: k = 3 kb + KR (KR is a template parameter so that every register
// index below is static).
template <int KR>
__device__ __forceinline__ void sweep_step(SolveSmem& sm, double (&a)[3][8], int kb, int rg, int cg, int hb) {
  constexpr int NR = (KR + 1) % 3;  // tile row of the next pivot
  const int k = 3 * kb + KR;
  const int cur = k & 1, nxt = cur ^ 1;
  __syncthreads();  // row k (published one step ahead) is visible
  double vcol[8];
  load_cols(sm.buf[cur], cg, vcol);
  const double* wr = &sm.wbuf[cur][3 * rg];
  double w[3];
  w[0] = wr[0];
  w[1] = wr[1];
  w[2] = wr[2];
  const bool piv_rg = (rg == kb);
  // look-ahead: the row of the next pivot is updated and published first
  const bool next_rg = (KR < 2) ? piv_rg : (rg == kb + 1);
  if (next_rg) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) a[NR][jj] = fma(w[NR], vcol[jj], a[NR][jj]);
    if (k + 1 < kN) publish_row(sm, a[NR], k + 1, cg, hb, nxt);
  }
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) {
    if (rr == NR) {
      if (!next_rg) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) a[rr][jj] = fma(w[rr], vcol[jj], a[rr][jj]);
      }
    } else if (rr == KR && piv_rg) {
      const double dinv = sm.piv[cur][0];
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

// Build K = c D P D + sigma I + A' diag(rho) A into the register tiles, then overwrite it
// with -K^-1 by the symmetric sweep operator.  Step k applies, with v = row k before the step,
//   a_rj <- a_rj - (v_r / d) v'_j   (r != k; v'_k = d - 1 makes the same update produce column k)
//   a_kj <- v_j / d,  a_kk <- -1/d
// and relies on a_rk == a_kr (symmetry) so the published ROW also supplies column k.
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
  if (rg == 0) publish_row(sm, a[0], 0, cg, hb, 0);
  for (int kb = 0; kb < kLegSteps; ++kb) {
    sweep_step<0>(sm, a, kb, rg, cg, hb);
    sweep_step<1>(sm, a, kb, rg, cg, hb);
    sweep_step<2>(sm, a, kb, rg, cg, hb);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__global__ void __launch_bounds__(kSolveThreads, 1)
admm_solve_kernel(const double* __restrict__ P_all, const double* __restrict__ q_all,
                  const float* __restrict__ l_all, const float* __restrict__ u_all,
                  const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                  float* __restrict__ x_all, int num, int* __restrict__ counter,
                  const __grid_constant__ SolveParams sp) {
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

    if (tid < kNP) {
      sm.Dp[tid] = (tid < kN) ? 1.0 : 0.0;
      sm.rhs[0][tid] = 0.0;
      sm.rhs[1][tid] = 0.0;
      sm.xD[tid] = 0.0;
      sm.buf[0][tid] = 0.0;
      sm.buf[1][tid] = 0.0;
      sm.wbuf[0][tid] = 0.0;
      sm.wbuf[1][tid] = 0.0;
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
    // ---- scaled data on the owning lanes ----
    const double c = sm.scal[0];
    const double cinv = 1.0 / c;
    const double Dinv = 1.0 / D, Einv = 1.0 / E;
    const double qb = c * D * q0;
    lb *= E;
    ub *= E;
    int ctype = 0;
    if (lb < -MPC_INFTY * 1e-4 && ub > MPC_INFTY * 1e-4) ctype = -1;
    else if (ub - lb < 1e-4) ctype = 1;
    double rho = sp.rho;
    double rv = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho : rho;
    double rinv = 1.0 / rv;
    // scaled constraint coefficients of the owned row: z~_i = cca * x~_lat + ccz * x~_z
    double cca, ccz;
    {
      const double Dlat = shfl(D, latsrc), Dz = shfl(D, zsrc);
      cca = (re == 4) ? 0.0 : E * Dlat;
      ccz = (re == 4) ? E * Dz : ((re & 1) ? -mu : mu) * E * Dz;
      if (!rown) { cca = 0.0; ccz = 0.0; }
    }
    auto build_G = [&]() {
      // A' diag(rho) A of this leg-step from the five row lanes
      const double r0 = rown ? rv : 0.0;
      const LegSums s1 = leg_reduce(r0 * cca * cca, r0 * ccz * ccz, hb);  // xx|yy on lanes 0|4, zz on lane 8
      const LegSums s2 = leg_reduce(r0 * cca * ccz, 0.0, hb);             // xz|yz on lanes 0|4
      if (cg == 0) { sm.G[rg * 6 + 0] = s1.lat; sm.G[rg * 6 + 1] = s2.lat; }
      if (cg == 4) { sm.G[rg * 6 + 2] = s1.lat; sm.G[rg * 6 + 3] = s2.lat; }
      if (cg == 8) sm.G[rg * 6 + 4] = s1.z;
    };
    build_G();
    if (vown) sm.rhs[1][vj] = -qb;  // rhs of iteration 1: x = z = y = 0
    __syncthreads();

    // ---- K3b: factor (explicit inverse in registers) ----
    factor_inverse(sm, a, rg, cg, hb, sigma);

    // ---- K4: ADMM iterations (osqp.c osqp_solve) ----
    double x = 0.0, z = 0.0, y = 0.0;  // x on variable lanes; z, y on row lanes
    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
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
      const double xt_lat = shfl(xt, latsrc), xt_z = shfl(xt, zsrc);
      const double zt = cca * xt_lat + ccz * xt_z;
      const double zr = alpha * zt + (1.0 - alpha) * z;
      double zn = zr + rinv * y;
      zn = fmin(fmax(zn, lb), ub);
      y = y + rv * (zr - zn);
      z = zn;
      // next rhs = sigma x - q + A'(rho z - y)
      {
        const double w = rown ? (rv * z - y) : 0.0;
        const LegSums s = leg_reduce(cca * w, ccz * w, hb);
        if (vown) sm.rhs[(iter + 1) & 1][vj] = sigma * x - qb + ((vc == 2) ? s.z : s.lat);
      }
      const bool can_check = sp.check_termination > 0 && (iter % sp.check_termination == 0);
      const bool can_adapt = sp.adaptive_rho && sp.adaptive_rho_interval > 0 &&
                             (iter % sp.adaptive_rho_interval == 0);
      const bool last = (iter == sp.max_iter);
      if (!(can_check || can_adapt || last)) continue;

      // ---- residuals (auxil.c compute_pri_res / compute_dua_res / tolerances) ----
      if (vown) sm.xD[vj] = D * x;
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      {
        const double x_lat = shfl(x, latsrc), x_z = shfl(x, zsrc);
        if (rown) {
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
          const double Px = c * D * sr;
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
      if (sm.flags[0]) {
        status = sm.flags[1];
        pri_res_out = sm.scal[4];
        break;
      }
      if (sm.flags[2]) {
        ++rho_updates;
        rho = sm.scal[2];
        rv = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho : rho;
        rinv = 1.0 / rv;
        // the rhs was built with the old rho vector: rebuild it, then refactor
        {
          const double w = rown ? (rv * z - y) : 0.0;
          const LegSums s = leg_reduce(cca * w, ccz * w, hb);
          if (vown) sm.rhs[(iter + 1) & 1][vj] = sigma * x - qb + ((vc == 2) ? s.z : s.lat);
        }
        build_G();
        __syncthreads();
        factor_inverse(sm, a, rg, cg, hb, sigma);
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    const double xo = D * x;
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
  }
}

}  // namespace mpcb200
