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
// Column of F_k in shared memory: (pos_c, vel_c) interleaved, so that the two columns a lane of the backward
// recursion needs are ONE 16-byte load per row (they were two 8-byte loads 48 bytes apart, two-way bank conflicts
// between the teams of a warp: half of the kernel's excess wavefronts)
__host__ __device__ constexpr int wrc_fcol(int j) { return j < 6 ? 2 * j : 2 * (j - 6) + 1; }

template <int H>
struct WrcSmem {
  static constexpr int n = 12 * H, m = 20 * H, nG = (H + 3) / 4;
  alignas(16) double Mt[H][6 * kMS];   // M~ = N^-1 G_ Delta^-1 (rows; first G_ Delta^-1 during a factorisation)
  alignas(16) double Fk[H][6 * kMS];   // F_k (rows; first G_ during a factorisation)
  alignas(16) double Nk[H][36];
  alignas(16) double Zk[H][36];
  alignas(16) double Phi[nG - 2][12 * kMS];  // Phi_1 .. Phi_(nG-2)
  alignas(16) double rhs[n];
  // iteration view:      pv | Xv | Pb | Xb | uv | ev   (recursion vectors, group boundaries, u, e / h)
  // factorisation view:  Lk [H][36] | Li [H][6]       (Cholesky factor of N_k and its inverse pivots)
  static constexpr int oPv = 0, oXv = oPv + (H + 1) * 12, oPb = oXv + (H + 1) * 12, oXb = oPb + (nG + 1) * 12,
                       oUv = oXb + (nG + 1) * 12, oEv = oUv + 6 * H, kItv = oEv + 6 * H;
  static_assert(kItv >= 42 * H, "the factorisation view must fit");
  alignas(16) double itv[kItv];
  // per leg-step constants of the iteration: kap[5] | Delta^-1 00 01 02 11 12 22 | q_ x y z  (read-only in the loop)
  alignas(16) double legc[4 * H][14];
  // per leg-step iterates of the five rows, normalised: zh[5] | uh[5]  (registers would starve the recursions)
  alignas(16) double zu[4 * H][10];
  alignas(16) double zero12[12];
  alignas(16) double Dp[n];
  // Ruiz: column-norm halves [2][n] | build: Q e [H][14] | factorisation: Riccati scratch | check: D x, G D x, S G D x
  alignas(16) double scr[2 * n];
  alignas(16) double B6t[3][12];       // top rows of B6c (step 0)
  double dT[9];                        // foot_drift: top rows of step k are B6t - k dT (same for every leg)
  double red[kWrcWarps * 16];
  double scal[16];                     // 0:c 1:1/c 2:rho 4:pri_res
  unsigned short be[H * H];            // sum_{i >= max(k,l)} (i - k)(i - l)  (< 2^16 for H <= 30)
  float st[48];
  int contacts[4 * H];
  int flags[8];                        // 0:done 1:status 2:refactor 3:problem index
};

// max_i |(G' S G)_ij| D_i over the rows i of two legs (6 hp .. 6 hp + 5 of every step), for column j of
// horizon step kj, component comp:  (G' S G)_ij = top_i . v + [comp_i == comp] vbm,  v = alpha u1 + beta u2.
// alpha and beta are the only things that depend on the block row, so the entry is  alpha A_i + beta B_i  with
// A_i = top_i . u1 + [.] vb1,  B_i = top_i . u2 + [.] vb2  formed once per item (with foot_drift the top rows are
// linear in the step, and so are A_i, B_i): three FP64 instructions per entry instead of five.
template <int H, bool kDrift>
__device__ __forceinline__ double wrc_colnorm_t(const WrcSmem<H>& sm, int hp, int kj, int comp, const double (&u1)[3],
                                                const double (&u2)[3], double vb1, double vb2, double dt2, double dt4) {
  double A0[6], B0[6], A1[6], B1[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double t0 = sm.B6t[0][6 * hp + i], t1 = sm.B6t[1][6 * hp + i], t2 = sm.B6t[2][6 * hp + i];
    const bool same = (i % 3) == comp;   // (6 hp + i) % 3 == i % 3
    A0[i] = fma(t0, u1[0], fma(t1, u1[1], fma(t2, u1[2], same ? vb1 : 0.0)));
    B0[i] = fma(t0, u2[0], fma(t1, u2[1], fma(t2, u2[2], same ? vb2 : 0.0)));
    if (kDrift) {
      const double d0 = sm.dT[i % 3], d1 = sm.dT[3 + i % 3], d2 = sm.dT[6 + i % 3];
      A1[i] = fma(d0, u1[0], fma(d1, u1[1], d2 * u1[2]));
      B1[i] = fma(d0, u2[0], fma(d1, u2[1], d2 * u2[2]));
    }
  }
  double mx = 0.0;
#pragma unroll 2
  for (int k = 0; k < H; ++k) {
    const int mxk = k > kj ? k : kj;
    const double a = (double)(H - mxk) * dt2, b = dt4 * (double)sm.be[H * k + kj];
    const double2* dk = reinterpret_cast<const double2*>(&sm.Dp[12 * k + 6 * hp]);
    const double2 d01 = dk[0], d23 = dk[1], d45 = dk[2];
    const double dd[6] = {d01.x, d01.y, d23.x, d23.y, d45.x, d45.y};
    const double kd = (double)k;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double Ai = kDrift ? fma(-kd, A1[i], A0[i]) : A0[i];
      const double Bi = kDrift ? fma(-kd, B1[i], B0[i]) : B0[i];
      const double e = fma(b, Bi, a * Ai);
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


// Horizon step of (warp, team).  Spread layout: a warp holds the steps at ONE position (warp / 2) of four groups
// (group = 4 (warp % 2) + team, step = 4 group + position), so that a round of a recursion sweep -- position r of
// every group -- is two warps working with all their teams, not eight warps working with one team each (the FP64
// pipe takes two cycles per warp instruction whatever the number of active lanes; with eight lanes of 32 active
// the sweeps of two co-resident CTAs were bound by it: 230 cycles per step against 150 alone).
__device__ __forceinline__ int wrc_stage(int warp, int team, int& pos, int& grp) {
  pos = warp >> 1;
  grp = 4 * (warp & 1) + team;
  return 4 * grp + pos;
}

// Shared-memory slot of a horizon step: position-major, so that the four teams of a warp (same position,
// consecutive groups) sit in CONSECUTIVE slots.  Indexed by the step itself they were four slots apart -- a multiple
// of 128 bytes for every per-step array -- and each of their loads hit the same banks four times over (ncu: 256 M
// bank conflicts per launch of 592 problems, LSU data pipe 50 % busy).  H = 30: 8 + 8 + 7 + 7 slots.
template <int H>
__device__ __forceinline__ int wrc_slot(int k) {
  constexpr int nG = (H + 3) / 4, full = H - 4 * (nG - 1);  // positions < full exist in the last group too
  const int pos = k & 3, grp = k >> 2;
  return (pos <= full ? nG * pos : nG * full + (nG - 1) * (pos - full)) + grp;
}

// The iterates a leg lane carries through a stretch of ADMM iterations.
struct WrcIter {
  double x[3];
};

#ifdef WRC_PROF
#define WRP(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); pc_[i] += t_ - *pm_; *pm_ = t_; } } while (0)
#else
#define WRP(i) do {} while (0)
#endif

// Shared-memory accesses of the iteration loop, by 32-bit shared address computed once per call.  The compiler, left
// alone at the 128-register cap, re-derived the addresses from %tid / %cluster_ctaid in every recursion step.  Every asm
// carries a "memory" clobber: without one NVVM merges identical `asm volatile` loads (seen in scripts/micro/
// lds_bench.cu) and may move loads across the loop's stores; the loop uses nothing but these for shared memory, so
// ordering against ordinary accesses is never in question.  (Same speed with and without the clobber: 135 k solves/s.)
__device__ __forceinline__ double2 wrc_ld2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double wrc_ld(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void wrc_st(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ uint32_t wrc_sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// `run` ADMM iterations.
#ifndef WRC_ITER_ATTR
#define WRC_ITER_ATTR __forceinline__
#endif
template <int H>
__device__ WRC_ITER_ATTR void wrc_iterate(WrcSmem<H>& sm, WrcIter& st, int run, double tzx, double tzy, double lo4,
                                         double hi4, double cyaw, double syaw, double dt, double sigma, double alpha
#ifdef WRC_PROF
                                         , long long* pc_, long long* pm_
#endif
                                         ) {
  using Smem = WrcSmem<H>;
  constexpr int nG = Smem::nG;
  constexpr uint32_t RS = kMS * 8;  // row stride of the 6 x 12 / 12 x 12 matrices in bytes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = lane >> 3, t8 = lane & 7;
  int pos, grp;
  const int kraw = wrc_stage(warp, team, pos, grp);
  const bool on = kraw < H;
  const int k = on ? kraw : H - 1;
  const bool isleg = on && t8 < 4, isax = on && t8 < 6;
  const int lg = t8 & 3;
  const int c = t8 < 6 ? t8 : t8 - 6;
  const int klast = (4 * grp + 3 < H - 1) ? 4 * grp + 3 : H - 1;
  const bool glast = (k == klast);
  // addresses
  const uint32_t a_itv = wrc_sa(sm.itv);
  const uint32_t a_pv = a_itv + 8 * Smem::oPv, a_Xv = a_itv + 8 * Smem::oXv, a_Pb = a_itv + 8 * Smem::oPb,
                 a_Xb = a_itv + 8 * Smem::oXb;
  const int sl = wrc_slot<H>(k);
  const uint32_t a_uk = a_itv + 8 * (Smem::oUv + 6 * sl), a_ek = a_itv + 8 * (Smem::oEv + 6 * sl);
  const uint32_t a_zero = wrc_sa(sm.zero12);
  const uint32_t a_mt = wrc_sa(&sm.Mt[sl][0]), a_fk = wrc_sa(&sm.Fk[sl][0]);
  const uint32_t a_nrow = wrc_sa(&sm.Nk[sl][6 * c]), a_zrow = wrc_sa(&sm.Zk[sl][6 * c]);
  const uint32_t a_rhsk = wrc_sa(&sm.rhs[12 * sl]), a_rhsj = a_rhsk + 24 * lg;
  const uint32_t a_lc = wrc_sa(&sm.legc[4 * sl + lg][0]);
  const uint32_t a_phi = wrc_sa(&sm.Phi[0][0]);
  // forward: reads X_k (group boundary for team 0), writes X_k+1; backward: reads p_k+1, writes p_k
  const int sln = wrc_slot<H>(glast ? k : k + 1);  // slot of the next step of the group
  const uint32_t xin = (pos == 0) ? a_Xb + 96 * grp : a_Xv + 96 * sl;
  const uint32_t xout = glast ? a_Xb + 96 * (grp + 1) : a_Xv + 96 * sln;
  const uint32_t pin = glast ? a_Pb + 96 * (grp + 1) : a_pv + 96 * sln;
  const uint32_t pout = (pos == 0) ? a_Pb + 96 * grp : a_pv + 96 * sl;
  const uint32_t xin1 = (pos == 0) ? a_zero : xin;             // first sweep: zero boundaries
  const uint32_t pin1 = glast ? a_zero : pin;
  const uint32_t pin3 = (glast && grp == nG - 1) ? a_zero : pin;   // p_H = 0 is the last group's true boundary
  // rotation coefficients of the lane's state pair (rows of Rt / Rt')
  const double fa = (c == 0) ? cyaw : (c == 1) ? -syaw : 0.0, fb = (c == 0) ? syaw : (c == 1) ? cyaw : 0.0;
  const double ba = (c == 0) ? cyaw : (c == 1) ? syaw : 0.0, bb = (c == 0) ? -syaw : (c == 1) ? cyaw : 0.0;
  const double fo = (c < 2) ? 0.0 : 1.0;
  const double oma = 1.0 - alpha;

  const uint32_t a_zu = wrc_sa(&sm.zu[4 * sl + lg][0]);
  double x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = st.x[i];

#pragma unroll 1
  for (int q_ = 0; q_ < run; ++q_) {
    // ---- u = M~ r (axis) ----
    __syncwarp();
    double uc;
    {
      double2 mv[6], rv[6];
#pragma unroll
      for (int h2 = 0; h2 < 6; ++h2) rv[h2] = wrc_ld2(a_rhsk + 16 * h2);
#pragma unroll
      for (int h2 = 0; h2 < 6; ++h2) mv[h2] = wrc_ld2(a_mt + RS * c + 16 * h2);
      double s0 = mv[0].x * rv[0].x, s1 = mv[0].y * rv[0].y, s2 = mv[1].x * rv[1].x, s3 = mv[1].y * rv[1].y;
#pragma unroll
      for (int h2 = 2; h2 < 6; h2 += 2) {
        s0 = fma(mv[h2].x, rv[h2].x, s0); s1 = fma(mv[h2].y, rv[h2].y, s1);
        s2 = fma(mv[h2 + 1].x, rv[h2 + 1].x, s2); s3 = fma(mv[h2 + 1].y, rv[h2 + 1].y, s3);
      }
      uc = (s0 + s1) + (s2 + s3);
      if (isax) wrc_st(a_uk + 8 * c, uc);
    }
    __syncwarp();
    // (N u)_c and t = -F' u (the addend of the backward recursion); F columns c and 6 + c stay in registers
    double fc[12], nuc, tpos, tvel;
    {
      const double2 u01 = wrc_ld2(a_uk), u23 = wrc_ld2(a_uk + 16), u45 = wrc_ld2(a_uk + 32);
      const double2 n01 = wrc_ld2(a_nrow), n23 = wrc_ld2(a_nrow + 16), n45 = wrc_ld2(a_nrow + 32);
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        const double2 f = wrc_ld2(a_fk + RS * d + 16 * c);   // F[d][c], F[d][6 + c]
        fc[d] = f.x;
        fc[6 + d] = f.y;
      }
      const double u6[6] = {u01.x, u01.y, u23.x, u23.y, u45.x, u45.y};
      nuc = fma(n45.x, u45.x, fma(n23.x, u23.x, n01.x * u01.x)) + fma(n45.y, u45.y, fma(n23.y, u23.y, n01.y * u01.y));
      double sp0 = fc[0] * u6[0], sp1 = fc[1] * u6[1], sv0 = fc[6] * u6[0], sv1 = fc[7] * u6[1];
#pragma unroll
      for (int d = 2; d < 6; d += 2) {
        sp0 = fma(fc[d], u6[d], sp0); sp1 = fma(fc[d + 1], u6[d + 1], sp1);
        sv0 = fma(fc[6 + d], u6[d], sv0); sv1 = fma(fc[7 + d], u6[d + 1], sv1);
      }
      tpos = -(sp0 + sp1); tvel = -(sv0 + sv1);
    }
    WRP(2);  // u, N u, t
    // ---- backward recursion  p_k = Acl_k' p_k+1 + t_k ----
    auto bstep = [&](uint32_t src, uint32_t dst, bool wr) {
      const double2 v01 = wrc_ld2(src + 48), v23 = wrc_ld2(src + 64), v45 = wrc_ld2(src + 80), p01 = wrc_ld2(src);
      const double own_pos = wrc_ld(src + 8 * c), own_vel = wrc_ld(src + 48 + 8 * c);
      double s0 = fma(fc[0], v01.x, tpos), s1 = fc[1] * v01.y, s2 = fma(fc[6], v01.x, tvel), s3 = fc[7] * v01.y;
      s0 = fma(fc[2], v23.x, s0); s1 = fma(fc[3], v23.y, s1);
      s2 = fma(fc[8], v23.x, s2); s3 = fma(fc[9], v23.y, s3);
      s0 = fma(fc[4], v45.x, s0); s1 = fma(fc[5], v45.y, s1);
      s2 = fma(fc[10], v45.x, s2); s3 = fma(fc[11], v45.y, s3);
      const double rot = fma(ba, p01.x, fma(bb, p01.y, fo * own_pos));
      const double np_ = own_pos + (s0 + s1);
      const double nv_ = fma(dt, rot, own_vel) + (s2 + s3);
      if (wr) { wrc_st(dst + 8 * c, np_); wrc_st(dst + 48 + 8 * c, nv_); }
    };
    // a round = position r of every group (two warps, all teams); the block barrier hands over to position r - 1
#pragma unroll 1
    for (int r_ = 3; r_ >= 0; --r_) {
      if (pos == r_) bstep(pin1, pout, isax);
      __syncthreads();
    }
    if (warp == 0) {
      const int i = lane < 12 ? lane : 0;
#pragma unroll 1
      for (int gj = nG - 2; gj >= 1; --gj) {
        const uint32_t aph = a_phi + (12 * kMS * 8) * (gj - 1) + 8 * i, apb = a_Pb + 96 * (gj + 1);
        double ph[12];
        double2 v[6];
#pragma unroll
        for (int h2 = 0; h2 < 6; ++h2) v[h2] = wrc_ld2(apb + 16 * h2);
        double s0 = wrc_ld(a_Pb + 96 * gj + 8 * i);
#pragma unroll
        for (int r = 0; r < 12; ++r) ph[r] = wrc_ld(aph + RS * r);
        double s1 = ph[1] * v[0].y, s2 = ph[2] * v[1].x, s3 = ph[3] * v[1].y;
        s0 = fma(ph[0], v[0].x, s0);
#pragma unroll
        for (int h2 = 2; h2 < 6; h2 += 2) {
          s0 = fma(ph[2 * h2], v[h2].x, s0); s1 = fma(ph[2 * h2 + 1], v[h2].y, s1);
          s2 = fma(ph[2 * h2 + 2], v[h2 + 1].x, s2); s3 = fma(ph[2 * h2 + 3], v[h2 + 1].y, s3);
        }
        if (lane < 12) wrc_st(a_Pb + 96 * gj + 8 * i, (s0 + s1) + (s2 + s3));
        __syncwarp();
      }
      // The warp that ran the boundary sweep re-reads its F columns: with them live across the sweep, the sweep's own
      // 24 operands no longer fit and ptxas issues its loads one by one.  Only this warp pays (a load instruction
      // costs the shared-memory pipe the same four cycles whatever its active lanes: scripts/micro/lds_bench.cu).
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        const double2 f = wrc_ld2(a_fk + RS * d + 16 * c);
        fc[d] = f.x;
        fc[6 + d] = f.y;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int r_ = 3; r_ >= 1; --r_) {
      if (pos == r_) bstep(pin3, pout, isax);
      __syncthreads();
    }
    WRP(3);  // backward sweep 3
    // ---- e = p_k+1,vel - u,  b = -Z e,  F row ----
    double fr[12], bfw, dl = 0.0;
    {
      const double ec = wrc_ld(pin3 + 48 + 8 * c) - uc;
      if (isax) wrc_st(a_ek + 8 * c, ec);
      const double2 z01 = wrc_ld2(a_zrow), z23 = wrc_ld2(a_zrow + 16), z45 = wrc_ld2(a_zrow + 32);
#pragma unroll
      for (int h2 = 0; h2 < 6; ++h2) {
        const double2 f = wrc_ld2(a_fk + RS * c + 16 * h2);   // F[c][h2], F[c][6 + h2]
        fr[h2] = f.x; fr[6 + h2] = f.y;
      }
      __syncwarp();
      const double2 e01 = wrc_ld2(a_ek), e23 = wrc_ld2(a_ek + 16), e45 = wrc_ld2(a_ek + 32);
      const double s0 = fma(z45.x, e45.x, fma(z23.x, e23.x, z01.x * e01.x));
      const double s1 = fma(z45.y, e45.y, fma(z23.y, e23.y, z01.y * e01.y));
      bfw = -(s0 + s1);
    }
    WRP(4);  // e, b, F row
    // ---- forward recursion  dlt_k = F_k X_k + b_k,  X_k+1 = A X_k + [0; dlt_k] ----
    auto fstep = [&](uint32_t src, uint32_t dst, bool wr) {
      double2 xv[6];
#pragma unroll
      for (int h2 = 0; h2 < 6; ++h2) xv[h2] = wrc_ld2(src + 16 * h2);
      const double own_pos = wrc_ld(src + 8 * c), own_vel = wrc_ld(src + 48 + 8 * c);
      double s0 = fma(fr[0], xv[0].x, bfw), s1 = fr[1] * xv[0].y, s2 = fr[2] * xv[1].x, s3 = fr[3] * xv[1].y;
#pragma unroll
      for (int h2 = 2; h2 < 6; h2 += 2) {
        s0 = fma(fr[2 * h2], xv[h2].x, s0); s1 = fma(fr[2 * h2 + 1], xv[h2].y, s1);
        s2 = fma(fr[2 * h2 + 2], xv[h2 + 1].x, s2); s3 = fma(fr[2 * h2 + 3], xv[h2 + 1].y, s3);
      }
      dl = (s0 + s1) + (s2 + s3);
      const double rot = fma(fa, xv[3].x, fma(fb, xv[3].y, fo * own_vel));
      if (wr) { wrc_st(dst + 8 * c, fma(dt, rot, own_pos)); wrc_st(dst + 48 + 8 * c, own_vel + dl); }
    };
#pragma unroll 1
    for (int r_ = 0; r_ < 4; ++r_) {
      if (pos == r_) fstep(xin1, xout, isax);
      __syncthreads();
    }
    if (warp == 0) {
      const int i = lane < 12 ? lane : 0;
#pragma unroll 1
      for (int gj = 1; gj <= nG - 2; ++gj) {
        const uint32_t aph = a_phi + (12 * kMS * 8) * (gj - 1) + RS * i, axb = a_Xb + 96 * gj;
        double2 v[6], f[6];
#pragma unroll
        for (int h2 = 0; h2 < 6; ++h2) v[h2] = wrc_ld2(axb + 16 * h2);
        double s0 = wrc_ld(a_Xb + 96 * (gj + 1) + 8 * i);
#pragma unroll
        for (int h2 = 0; h2 < 6; ++h2) f[h2] = wrc_ld2(aph + 16 * h2);
        double s1 = f[0].y * v[0].y, s2 = f[1].x * v[1].x, s3 = f[1].y * v[1].y;
        s0 = fma(f[0].x, v[0].x, s0);
#pragma unroll
        for (int h2 = 2; h2 < 6; h2 += 2) {
          s0 = fma(f[h2].x, v[h2].x, s0); s1 = fma(f[h2].y, v[h2].y, s1);
          s2 = fma(f[h2 + 1].x, v[h2 + 1].x, s2); s3 = fma(f[h2 + 1].y, v[h2 + 1].y, s3);
        }
        if (lane < 12) wrc_st(a_Xb + 96 * (gj + 1) + 8 * i, (s0 + s1) + (s2 + s3));
        __syncwarp();
      }
#pragma unroll
      for (int h2 = 0; h2 < 6; ++h2) {
        const double2 f = wrc_ld2(a_fk + RS * c + 16 * h2);
        fr[h2] = f.x; fr[6 + h2] = f.y;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int r_ = 0; r_ < 3; ++r_) {
      if (pos == r_) fstep(xin, xout, isax && !glast);
      __syncthreads();
    }
    if (pos == 3) fstep(xin, xout, false);
    WRP(5);  // forward recursion
    // ---- h = N u - dlt (axis);  x~ = a - M~' h, row updates, next rhs (leg) ----
    if (isax) wrc_st(a_ek + 8 * c, nuc - dl);
    {
      // everything that does not wait for h first: the leg's constant block, its rhs, its three columns of M~
      double2 l2[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) l2[i] = wrc_ld2(a_lc + 16 * i);  // kap0 kap1 | kap2 kap3 | kap4 di0 | di1 di2 | di3 di4 | di5 qx | qy qz
      const double rh0 = wrc_ld(a_rhsj), rh1 = wrc_ld(a_rhsj + 8), rh2 = wrc_ld(a_rhsj + 16);
      double z[5], u[5];
      {
        const double2 z01 = wrc_ld2(a_zu), z23 = wrc_ld2(a_zu + 16), z4u0 = wrc_ld2(a_zu + 32), u12 = wrc_ld2(a_zu + 48),
                      u34 = wrc_ld2(a_zu + 64);
        z[0] = z01.x; z[1] = z01.y; z[2] = z23.x; z[3] = z23.y; z[4] = z4u0.x;
        u[0] = z4u0.y; u[1] = u12.x; u[2] = u12.y; u[3] = u34.x; u[4] = u34.y;
      }
      double mc[6][3];
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        mc[cc][0] = wrc_ld(a_mt + RS * cc + 24 * lg);
        mc[cc][1] = wrc_ld(a_mt + RS * cc + 24 * lg + 8);
        mc[cc][2] = wrc_ld(a_mt + RS * cc + 24 * lg + 16);
      }
      const double a0 = l2[2].y * rh0 + l2[3].x * rh1 + l2[3].y * rh2;   // a = Delta^-1 r
      const double a1 = l2[3].x * rh0 + l2[4].x * rh1 + l2[4].y * rh2;
      const double a2 = l2[3].y * rh0 + l2[4].y * rh1 + l2[5].x * rh2;
      __syncwarp();
      const double2 h01 = wrc_ld2(a_ek), h23 = wrc_ld2(a_ek + 16), h45 = wrc_ld2(a_ek + 32);
      const double h6[6] = {h01.x, h01.y, h23.x, h23.y, h45.x, h45.y};
      double sx0 = mc[0][0] * h6[0], sx1 = mc[1][0] * h6[1], sy0 = mc[0][1] * h6[0], sy1 = mc[1][1] * h6[1],
             sz0 = mc[0][2] * h6[0], sz1 = mc[1][2] * h6[1];
#pragma unroll
      for (int cc = 2; cc < 6; cc += 2) {
        sx0 = fma(mc[cc][0], h6[cc], sx0); sx1 = fma(mc[cc + 1][0], h6[cc + 1], sx1);
        sy0 = fma(mc[cc][1], h6[cc], sy0); sy1 = fma(mc[cc + 1][1], h6[cc + 1], sy1);
        sz0 = fma(mc[cc][2], h6[cc], sz0); sz1 = fma(mc[cc + 1][2], h6[cc + 1], sz1);
      }
      const double xtx = a0 - (sx0 + sx1), xty = a1 - (sy0 + sy1), xtz = a2 - (sz0 + sz1);
      x[0] = alpha * xtx + oma * x[0];
      x[1] = alpha * xty + oma * x[1];
      x[2] = alpha * xtz + oma * x[2];
      const double zt[5] = {fma(tzx, xtz, xtx), fma(-tzx, xtz, xtx), fma(tzy, xtz, xty), fma(-tzy, xtz, xty), xtz};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const double zr = alpha * zt[i] + oma * z[i];
        double zn = zr + u[i];
        if (i == 0 || i == 2) zn = (zn < 0.0) ? 0.0 : zn;
        else if (i == 1 || i == 3) zn = (zn > 0.0) ? 0.0 : zn;
        else { zn = (zn < lo4) ? lo4 : zn; zn = (zn > hi4) ? hi4 : zn; }
        u[i] = u[i] + (zr - zn);
        z[i] = zn;
      }
      // next rhs = sigma x - q + A_'(rho z - y)
      const double e0 = l2[0].x * (z[0] - u[0]), e1 = l2[0].y * (z[1] - u[1]), e2 = l2[1].x * (z[2] - u[2]),
                   e3 = l2[1].y * (z[3] - u[3]), e4 = l2[2].x * (z[4] - u[4]);
      const double r0_ = sigma * x[0] - l2[5].y + (e0 + e1);
      const double r1_ = sigma * x[1] - l2[6].x + (e2 + e3);
      const double r2_ = sigma * x[2] - l2[6].y + (e4 + (tzx * (e0 - e1) + tzy * (e2 - e3)));
      if (isleg) {
        wrc_st(a_rhsj, r0_); wrc_st(a_rhsj + 8, r1_); wrc_st(a_rhsj + 16, r2_);
#pragma unroll
        for (int i = 0; i < 5; ++i) { wrc_st(a_zu + 8 * i, z[i]); wrc_st(a_zu + 40 + 8 * i, u[i]); }
      }
    }
    WRP(6);  // x~, rows, rhs
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 3; ++i) st.x[i] = x[i];
}


template <int H>
__global__ void __launch_bounds__(kWrcThreads, kWrcCtasPerSm)
wrench_riccati_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait,
                      MpcResult* __restrict__ results, float* __restrict__ x_all, int num, int* __restrict__ counter,
                      double* __restrict__ warm, int warm_stride, double* __restrict__ gscr,
                      const MpcTorqueIn* __restrict__ tin,
                      MpcTorqueOut* __restrict__ tout, const __grid_constant__ BuildParams bp,
                      const __grid_constant__ SolveParams sp) {
  using Smem = WrcSmem<H>;
  constexpr int n = Smem::n, m = Smem::m, nG = Smem::nG;
  constexpr int kWX = 0, kWQ = n, kWZ = 2 * n, kWY = 2 * n + m, kWRho = 2 * n + 2 * m, kWLive = kWRho + 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  double* const s_pv = sm.itv + Smem::oPv;
  double* const s_Xv = sm.itv + Smem::oXv;
  double* const s_Pb = sm.itv + Smem::oPb;
  double* const s_Xb = sm.itv + Smem::oXb;
  double* const s_uv = sm.itv + Smem::oUv;
  double* const s_ev = sm.itv + Smem::oEv;
  double (*const s_Lk)[36] = reinterpret_cast<double (*)[36]>(sm.itv);
  double (*const s_Li)[6] = reinterpret_cast<double (*)[6]>(sm.itv + 36 * H);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = lane >> 3, t8 = lane & 7;
  int pos_, grp_;
  const int kraw = wrc_stage(warp, team, pos_, grp_);
  const bool on = kraw < H;
  const int k = on ? kraw : H - 1;                      // idle teams shadow the last step (never write)
  const bool isleg = on && t8 < 4, isax = on && t8 < 6;
  const int lg = t8 & 3;                                // leg (leg role)
  const int c = t8 < 6 ? t8 : t8 - 6;                   // component (axis role)
  const int j0 = 12 * k + 3 * lg, r0 = 5 * (4 * k + lg);  // first variable / first row of the leg-step
  const int sl = wrc_slot<H>(k);                           // shared-memory slot of the step's matrices and vectors
  const int js = 12 * sl + 3 * lg;                         // the leg's entries of rhs

  const bool kWarm = warm != nullptr;
  const double mu = sp.mu, sigma = sp.sigma, alpha = sp.alpha;
  const double dt = bp.dt, inv_m = 1.0 / bp.mass;
  const double dt2 = dt * dt, dt4 = dt2 * dt2;

  // ---- once per CTA ----
  for (int i = tid; i < H * H; i += kWrcThreads) {
    const int kk = i / H, ll = i - H * kk;
    const int mxk = kk > ll ? kk : ll;
    int s = 0;
    for (int q = mxk; q < H; ++q) s += (q - kk) * (q - ll);
    sm.be[i] = (unsigned short)s;
  }
  if (tid < 12) sm.zero12[tid] = 0.0;

  // Profiling build only (-DWRC_PROF; scripts/prof_wrc.py): thread 0 accumulates clock64 per phase and prints
  // the totals of the first two problems.  Compiled out otherwise.
#undef WRP
#ifdef WRC_PROF
  long long pc_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pm_ = 0;
#define WRP(i) do { if (tid == 0) { const long long t_ = clock64(); pc_[i] += t_ - pm_; pm_ = t_; } } while (0)
#else
#define WRP(i) do {} while (0)
#endif
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;
#ifdef WRC_PROF
    if (tid == 0) pm_ = clock64();
#endif

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
    double* const gamv = sm.scr + 14 * H;  // gam (6H), next to Q e; both are dead once the gradient is formed
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
      gamv[tid] = s;
    }
    __syncthreads();

    // ---- per-leg constants of the build (leg role) ----
    const bool drift = bp.foot_drift != 0;
    // top[i][q]: row i of B6c, column 3 lg + q, at this step (re-formed where it is needed: nine registers less
    // across the iteration loop)
    auto load_top = [&](double (&top)[3][3]) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) top[i][q] = fma(-(double)k, sm.dT[3 * i + q], sm.B6t[i][3 * lg + q]);
    };
    // gradient (ConvexMpc.cpp:215-217): q_j = B6c[:, j] . gam_k
    auto gradient = [&](const double (&top)[3][3], double (&q0v)[3]) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        q0v[q] = top[0][q] * gamv[6 * k] + top[1][q] * gamv[6 * k + 1] + top[2][q] * gamv[6 * k + 2] +
                 inv_m * gamv[6 * k + 3 + q];
    };
    double qb[3], pjj[3], qsc[3];
    {
      double top[3][3], q0v[3];
      load_top(top);
      gradient(top, q0v);
      const double a = (double)(H - k) * dt2, b = dt4 * (double)sm.be[H * k + k];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        qb[q] = q0v[q];  // scaled after the equilibration
        qsc[q] = (live && isleg) ? ws[kWQ + j0 + q] : q0v[q];
        if (kWarm && isleg) ws[kWQ + j0 + q] = q0v[q];  // the slot's q of the NEXT update (the old one was just read)
        // diagonal entry of P (the only one R2 touches): P_jj = B6c_j' S_kk B6c_j + 2 r_j
        const double u10 = bp.Qd[6] * top[0][q], u11 = bp.Qd[7] * top[1][q], u12 = bp.Qd[8] * top[2][q];
        const double u20 = th00 * top[0][q] + th01 * top[1][q], u21 = th01 * top[0][q] + th11 * top[1][q],
                     u22 = th22 * top[2][q];
        const double vb1 = bp.Qd[9 + q] * inv_m * inv_m, vb2 = bp.Qd[3 + q] * inv_m * inv_m;
        pjj[q] = top[0][q] * (a * u10 + b * u20) + top[1][q] * (a * u11 + b * u21) + top[2][q] * (a * u12 + b * u22) +
                 (a * vb1 + b * vb2) + bp.Rd[3 * lg + q];
      }
    }
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
    WRP(0);  // load, build, Ruiz
    if (tid == 0) { sm.scal[0] = cs; sm.scal[1] = 1.0 / cs; }
#pragma unroll
    for (int q = 0; q < 3; ++q) qb[q] = cs * D[q] * qb[q];
    // constraint types (auxil.c set_rho_vec): -1 loose, 1 equality, 0 inequality
    auto ctype_of = [](double lo, double hi) {
      return (lo < -MPC_INFTY * 1e-4 && hi > MPC_INFTY * 1e-4) ? -1 : ((hi - lo < 1e-4) ? 1 : 0);
    };
    auto rho_of = [](int ct, double rho) { return (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho; };
    // (two bits per row, value + 1; D, cca = E_row D_own and the rho vector are re-read / re-derived where they
    // are needed instead of living in registers across the iteration loop)
    int ctp;
    {
      const int c0 = ctype_of(0.0, (double)(float)MPC_INFTY * E[0]), c1 = ctype_of(-(double)(float)MPC_INFTY * E[1], 0.0);
      const int c2 = ctype_of(0.0, (double)(float)MPC_INFTY * E[2]), c3 = ctype_of(-(double)(float)MPC_INFTY * E[3], 0.0);
      const int c4 = ctype_of(lo4 * E[4], hi4 * E[4]);
      ctp = (c0 + 1) | ((c1 + 1) << 2) | ((c2 + 1) << 4) | ((c3 + 1) << 6) | ((c4 + 1) << 8);
    }
    auto rho_row = [&](int i, double rho) { return rho_of(((ctp >> (2 * i)) & 3) - 1, rho); };
    // Rows in normalised form (wrench_kernel.cuh): row i = cca_i (x_own +- t x_fz), the kernel iterates on
    // zh = z / cca and uh = y / (rho cca);  kap = rho cca^2.
    double* const lc = sm.legc[4 * sl + lg];
    double kap[5];
    const double tzx = mu * D[2] / D[0], tzy = mu * D[2] / D[1];
    // iterates: x in registers, zh / uh of the five rows in shared memory (zu); cca = E_row D_own of the rows goes
    // to the CTA's global scratch line (L2): only the residual checks, rho updates and the warm slot read it
    double x[3] = {0.0, 0.0, 0.0};
    double* const zu = sm.zu[4 * sl + lg];
    double* const gcca = gscr + size_t(blockIdx.x) * (20 * H) + r0;
    {
      const double cca[5] = {E[0] * D[0], E[1] * D[0], E[2] * D[1], E[3] * D[1], E[4] * D[2]};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const double rvi = rho_row(i, rho0);
        kap[i] = rvi * cca[i] * cca[i];
        if (isleg) {
          gcca[i] = cca[i];
          zu[i] = live ? ws[kWZ + r0 + i] / cca[i] : 0.0;
          zu[5 + i] = live ? ws[kWY + r0 + i] / (rvi * cca[i]) : 0.0;
        }
      }
      lo4 = lo4 * E[4] / cca[4];
      hi4 = hi4 * E[4] / cca[4];
      if (live && isleg) {
#pragma unroll
        for (int q = 0; q < 3; ++q) x[q] = ws[kWX + j0 + q];
      }
    }
    double rho_cur = rho0;
    if (isleg) {
#pragma unroll
      for (int i = 0; i < 5; ++i) lc[i] = kap[i];
#pragma unroll
      for (int q = 0; q < 3; ++q) lc[11 + q] = qb[q];
    }

    // rhs = sigma x - q + A_'(rho z - y)   (kap and q_ from the leg's constant block)
    auto publish_rhs = [&]() {
      const double2* l2 = reinterpret_cast<const double2*>(lc);
      const double2 k01 = l2[0], k23 = l2[1], k4_ = l2[2], dq = l2[5], q12 = l2[6];
      const double e0 = k01.x * (zu[0] - zu[5]), e1 = k01.y * (zu[1] - zu[6]), e2 = k23.x * (zu[2] - zu[7]),
                   e3 = k23.y * (zu[3] - zu[8]), e4 = k4_.x * (zu[4] - zu[9]);
      const double r0_ = sigma * x[0] - dq.y + (e0 + e1);
      const double r1_ = sigma * x[1] - q12.x + (e2 + e3);
      const double r2_ = sigma * x[2] - q12.y + (e4 + (tzx * (e0 - e1) + tzy * (e2 - e3)));
      if (isleg) { sm.rhs[js] = r0_; sm.rhs[js + 1] = r1_; sm.rhs[js + 2] = r2_; }
    };

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
        const double D0 = sm.Dp[j0], D1 = sm.Dp[j0 + 1], D2 = sm.Dp[j0 + 2];
        double di[6];
        {
          const double kap[5] = {lc[0], lc[1], lc[2], lc[3], lc[4]};
          const double r2x = bp.Rd[3 * lg], r2y = bp.Rd[3 * lg + 1], r2z = bp.Rd[3 * lg + 2];
          const double sa = kap[0] + kap[1], sb = kap[2] + kap[3];
          const double dxx = cs * D0 * D0 * r2x + sigma + sa, dyy = cs * D1 * D1 * r2y + sigma + sb;
          const double dzz = cs * D2 * D2 * r2z + sigma + kap[4] + tzx * tzx * sa + tzy * tzy * sb;
          const double dxz = tzx * (kap[0] - kap[1]), dyz = tzy * (kap[2] - kap[3]);
          const double m00 = dyy * dzz - dyz * dyz, m11 = dxx * dzz - dxz * dxz, m22 = dxx * dyy;
          const double idet = 1.0 / (dxx * m00 - dxz * dxz * dyy);
          di[0] = m00 * idet; di[1] = dxz * dyz * idet; di[2] = -dyy * dxz * idet;
          di[3] = m11 * idet; di[4] = -dxx * dyz * idet; di[5] = m22 * idet;
          if (isleg) {
#pragma unroll
            for (int i = 0; i < 6; ++i) lc[5 + i] = di[i];
          }
        }
        // G_ (into Fk) and M1 = G_ Delta^-1 (into Mt): the leg's three columns.  G_ = dt B6c D: the step's
        // velocity increment per unit of (scaled) force
        if (isleg) {
          const double dm[3][3] = {{di[0], di[1], di[2]}, {di[1], di[3], di[4]}, {di[2], di[4], di[5]}};
          const double Dq[3] = {D0, D1, D2};
          double top[3][3];
          load_top(top);
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double gq[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const double b6 = (cc < 3) ? top[cc < 3 ? cc : 0][q] : ((cc - 3 == q) ? inv_m : 0.0);
              gq[q] = dt * Dq[q] * b6;
              sm.Fk[sl][cc * kMS + 3 * lg + q] = gq[q];
            }
#pragma unroll
            for (int q = 0; q < 3; ++q)
              sm.Mt[sl][cc * kMS + 3 * lg + q] = gq[0] * dm[0][q] + gq[1] * dm[1][q] + gq[2] * dm[2][q];
          }
        }
        __syncwarp();
        // N = M1 G_' (row c per axis lane)
        if (isax) {
          double mr[12];
#pragma unroll
          for (int h2 = 0; h2 < 6; ++h2) {
            const double2 v = *reinterpret_cast<const double2*>(&sm.Mt[sl][c * kMS + 2 * h2]);
            mr[2 * h2] = v.x; mr[2 * h2 + 1] = v.y;
          }
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            double s = 0.0;
#pragma unroll
            for (int h2 = 0; h2 < 6; ++h2) {
              const double2 v = *reinterpret_cast<const double2*>(&sm.Fk[sl][d * kMS + 2 * h2]);
              s = fma(mr[2 * h2], v.x, s);
              s = fma(mr[2 * h2 + 1], v.y, s);
            }
            sm.Nk[sl][6 * c + d] = s;
          }
        }
        __syncwarp();
        // Cholesky N_k = L L' (one lane per step; a non-positive pivot zeroes its column)
        if (on && t8 == 0) {
          double Lm[6][6];
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Lm[a][b] = sm.Nk[sl][6 * a + b];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double d = Lm[cc][cc];
#pragma unroll
            for (int q = 0; q < cc; ++q) d -= Lm[cc][q] * Lm[cc][q];
            const bool ok = d > 0.0;
            const double ld = ok ? sqrt(d) : 0.0, li = ok ? 1.0 / ld : 0.0;
            Lm[cc][cc] = ld;
            s_Li[k][cc] = li;
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
            for (int b = 0; b < 6; ++b) s_Lk[k][6 * a + b] = (b <= a) ? Lm[a][b] : 0.0;
        }
        __syncwarp();
        // M~ = L^-T L^-1 M1: the leg's three columns, forward then backward substitution
        if (isleg) {
          const double* Lp = s_Lk[k];
          const double* li = s_Li[k];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            double y6[6];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
              double s = sm.Mt[sl][cc * kMS + 3 * lg + q];
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
            for (int cc = 0; cc < 6; ++cc) sm.Mt[sl][cc * kMS + 3 * lg + q] = y6[cc];
          }
        }
        __syncthreads();
        WRP(8);  // factorisation: per-step part (Delta^-1, G_, N, Cholesky, M~)
        // Riccati recursion on warp 0 (scratch in scr): Pi 144 | G 144 | U 72 | T1 36 | Mi 36.  Six short phases per
        // step with fixed lane roles and unrolled index arithmetic:
        //   A  rows of Y = Pi A in registers -> U = Y_v, G = A' Y (12 lanes), T1 = Pi_vv L (6 lanes)
        //   B  Mm = I + L' T1, row per lane;  C  Mm^-1 by Gauss-Jordan through shuffles (pivots >= 1)
        //   D  Z = L Mm^-1 L' (21 lanes, upper triangle mirrored);  E  F = -Z U (24 lanes);
        //   F  Pi <- cQ + sym(G) + U' F (12 lanes, entries j >= i mirrored)
        // (Spreading the phases over the whole CTA, one output per thread and a block barrier between phases, was
        // measured too: 118 k solves/s against 122 k -- the other CTA of the SM fills the idle warps' slots anyway.)
        if (warp == 0) {
          double* const Pi = sm.scr;
          double* const G = sm.scr + 144;
          double* const U = sm.scr + 288;
          double* const T1 = sm.scr + 360;
          double* const Mi = sm.scr + 396;
          for (int i = lane; i < 144; i += 32) Pi[i] = (i / 12 == i % 12) ? cs * bp.Qd[i / 12] : 0.0;
          __syncwarp();
          // (v R~)[b] of a 6-vector
          auto colR = [&](const double (&v)[6], double (&o)[6]) {
            o[0] = cyaw * v[0] - syaw * v[1];
            o[1] = syaw * v[0] + cyaw * v[1];
            o[2] = v[2]; o[3] = v[3]; o[4] = v[4]; o[5] = v[5];
          };
          // upper-triangle index of lane < 21 for phase D
          int za = 0, zb = lane < 21 ? lane : 0;
          while (zb >= 6 - za) { zb -= 6 - za; ++za; }
          zb += za;
#pragma unroll 1
          for (int ks = H - 1; ks >= 0; --ks) {
            const double* Lp = s_Lk[ks];
            const int ss = wrc_slot<H>(ks);
            // ---- A ----
            if (lane < 12) {
              const int i = lane;
              double yr[12];  // row i of Y = Pi A: columns 0-5 unchanged, columns 6-11 gain dt (Pi[i, 0:6] R~)
              {
                double pr[12], t6[6];
#pragma unroll
                for (int h2 = 0; h2 < 6; ++h2) {
                  const double2 v = *reinterpret_cast<const double2*>(&Pi[12 * i + 2 * h2]);
                  pr[2 * h2] = v.x; pr[2 * h2 + 1] = v.y;
                }
                const double p6[6] = {pr[0], pr[1], pr[2], pr[3], pr[4], pr[5]};
                colR(p6, t6);
#pragma unroll
                for (int j = 0; j < 6; ++j) { yr[j] = pr[j]; yr[6 + j] = fma(dt, t6[j], pr[6 + j]); }
              }
              if (i < 6) {
#pragma unroll
                for (int j = 0; j < 12; ++j) G[12 * i + j] = yr[j];          // G rows 0-5 = Y rows 0-5
              } else {
                // G row 6 + a = Y row 6 + a + dt (R~' Y[0:6, :])[a, :]; the pos rows of Y it needs are re-formed here
                const int a6 = i - 6;
                const int ra = a6 < 2 ? 0 : a6, rb = a6 < 2 ? 1 : a6;
                const double ca = a6 == 0 ? cyaw : a6 == 1 ? syaw : 1.0, cb = a6 == 0 ? -syaw : a6 == 1 ? cyaw : 0.0;
                double ya[12], yb[12];
                {
                  double pa[12], pb[12], ta[6], tb[6];
#pragma unroll
                  for (int h2 = 0; h2 < 6; ++h2) {
                    const double2 v = *reinterpret_cast<const double2*>(&Pi[12 * ra + 2 * h2]);
                    const double2 w = *reinterpret_cast<const double2*>(&Pi[12 * rb + 2 * h2]);
                    pa[2 * h2] = v.x; pa[2 * h2 + 1] = v.y; pb[2 * h2] = w.x; pb[2 * h2 + 1] = w.y;
                  }
                  const double a6v[6] = {pa[0], pa[1], pa[2], pa[3], pa[4], pa[5]};
                  const double b6v[6] = {pb[0], pb[1], pb[2], pb[3], pb[4], pb[5]};
                  colR(a6v, ta);
                  colR(b6v, tb);
#pragma unroll
                  for (int j = 0; j < 6; ++j) {
                    ya[j] = pa[j]; ya[6 + j] = fma(dt, ta[j], pa[6 + j]);
                    yb[j] = pb[j]; yb[6 + j] = fma(dt, tb[j], pb[6 + j]);
                  }
                }
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                  U[12 * a6 + j] = yr[j];
                  G[12 * i + j] = fma(dt, fma(ca, ya[j], cb * yb[j]), yr[j]);
                }
              }
            } else if (lane < 18) {
              // T1 = Pi_vv L, row a
              const int a = lane - 12;
              double pv6[6];
#pragma unroll
              for (int q = 0; q < 6; ++q) pv6[q] = Pi[12 * (6 + a) + 6 + q];
#pragma unroll
              for (int b = 0; b < 6; ++b) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q)
                  if (q >= b) s = fma(pv6[q], Lp[6 * q + b], s);
                T1[6 * a + b] = s;
              }
            }
            __syncwarp();
            // ---- B, C ----
            {
              const int r = lane < 6 ? lane : 0;
              double row[6];
#pragma unroll
              for (int b = 0; b < 6; ++b) {
                double s = (r == b) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) s = fma(Lp[6 * q + r], T1[6 * q + b], s);   // L[q][r] = 0 for q < r
                row[b] = s;
              }
#pragma unroll
              for (int pv_ = 0; pv_ < 6; ++pv_) {
                double rowp[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) rowp[j] = __shfl_sync(0xffffffffu, row[j], pv_);
                // reciprocal of the pivot (>= 1): fp32 seed, two Newton steps in fp64 (a third of a division's latency)
                double d = (double)__frcp_rn((float)rowp[pv_]);
                d = fma(d, fma(-rowp[pv_], d, 1.0), d);
                d = fma(d, fma(-rowp[pv_], d, 1.0), d);
                const bool piv = (lane == pv_);
                const double f = row[pv_] * d;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                  const double upd = (j == pv_) ? -f : row[j] - f * rowp[j];
                  const double prw = (j == pv_) ? d : rowp[j] * d;
                  row[j] = piv ? prw : upd;
                }
              }
              if (lane < 6) {
#pragma unroll
                for (int j = 0; j < 6; ++j) Mi[6 * lane + j] = row[j];
              }
            }
            __syncwarp();
            // ---- D: Z[za][zb] = sum_p L[za][p] (sum_q Mi[p][q] L[zb][q]) ----
            if (lane < 21) {
              double lb6[6], la6[6];
#pragma unroll
              for (int q = 0; q < 6; ++q) { lb6[q] = Lp[6 * zb + q]; la6[q] = Lp[6 * za + q]; }
              double s = 0.0;
#pragma unroll
              for (int p_ = 0; p_ < 6; ++p_) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) t = fma(Mi[6 * p_ + q], lb6[q], t);
                s = fma(la6[p_], t, s);
              }
              sm.Zk[ss][6 * za + zb] = s;
              sm.Zk[ss][6 * zb + za] = s;
            }
            __syncwarp();
            // ---- E: F = -Z U, lane (a, three columns) ----
            if (lane < 24) {
              const int a = lane % 6, j3 = 3 * (lane / 6);
              double zr[6];
#pragma unroll
              for (int q = 0; q < 6; ++q) zr[q] = sm.Zk[ss][6 * a + q];
#pragma unroll
              for (int jj = 0; jj < 3; ++jj) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) s = fma(zr[q], U[12 * q + j3 + jj], s);
                sm.Fk[ss][a * kMS + wrc_fcol(j3 + jj)] = -s;
              }
            }
            __syncwarp();
            // ---- F: Pi <- cQ + sym(G) + U' F, entries j >= i of row i, mirrored ----
            if (lane < 12) {
              const int i = lane;
              double ui[6];
#pragma unroll
              for (int q = 0; q < 6; ++q) ui[q] = U[12 * q + i];
#pragma unroll
              for (int j = 0; j < 12; ++j) {
                if (j >= i) {
                  double s = 0.5 * (G[12 * i + j] + G[12 * j + i]);
#pragma unroll
                  for (int q = 0; q < 6; ++q) s = fma(ui[q], sm.Fk[ss][q * kMS + wrc_fcol(j)], s);
                  if (i == j) s += cs * bp.Qd[i];
                  Pi[12 * i + j] = s;
                  Pi[12 * j + i] = s;
                }
              }
            }
            __syncwarp();
          }
        }
        __syncthreads();
        WRP(9);  // factorisation: Riccati recursion
        // group transitions Phi_j = Acl_(4j+3) ... Acl_4j, j = 1 .. nG-2: one thread per column
        if (tid < (nG - 2) * 12) {
          const int gj = 1 + tid / 12, col = tid % 12;
          double T[12], Tn[6];
#pragma unroll
          for (int i = 0; i < 12; ++i) T[i] = (i == col) ? 1.0 : 0.0;
#pragma unroll 1
          for (int s_ = 0; s_ < 4; ++s_) {
            const double* Fp = sm.Fk[wrc_slot<H>(4 * gj + s_)];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
              double acc = T[6 + a];
#pragma unroll
              for (int i = 0; i < 12; ++i) acc = fma(Fp[a * kMS + wrc_fcol(i)], T[i], acc);
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
        if (tid >= 128 && tid < 140) s_Xb[tid - 128] = 0.0;  // X_0 = 0: the true boundary of the first group
        publish_rhs();
        __syncthreads();
        WRP(1);  // factorisation
      }
      int run = until_check < until_adapt ? until_check : until_adapt;
      run = run < sp.max_iter - iter ? run : sp.max_iter - iter;
      {
        WrcIter st_;
#pragma unroll
        for (int i = 0; i < 3; ++i) st_.x[i] = x[i];
#ifdef WRC_PROF
        wrc_iterate<H>(sm, st_, run, tzx, tzy, lo4, hi4, cyaw, syaw, dt, sigma, alpha, pc_, &pm_);
#else
        wrc_iterate<H>(sm, st_, run, tzx, tzy, lo4, hi4, cyaw, syaw, dt, sigma, alpha);
#endif
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = st_.x[i];
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
      const double D[3] = {sm.Dp[j0], sm.Dp[j0 + 1], sm.Dp[j0 + 2]};
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
      // (maxima of non-negative values through their bit patterns: a NaN -- a bad state record -- compares above
      // everything and reaches the test below, where it fails every comparison; fmax would drop it and report
      // the problem solved at the first check)
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      if (isleg) {
        const double cca[5] = {gcca[0], gcca[1], gcca[2], gcca[3], gcca[4]};
        const double z[5] = {zu[0], zu[1], zu[2], zu[3], zu[4]}, u[5] = {zu[5], zu[6], zu[7], zu[8], zu[9]};
        double top[3][3];
        load_top(top);
        const double Ah[5] = {fma(tzx, x[2], x[0]), fma(-tzx, x[2], x[0]), fma(tzy, x[2], x[1]), fma(-tzy, x[2], x[1]), x[2]};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const double Dn = (i < 2) ? D[0] : (i < 4) ? D[1] : D[2];
          const double rp = Ah[i] - z[i];
          v[0] = max_bits(v[0], fabs(cca[i] * rp));   // scaled primal residual
          v[1] = max_bits(v[1], Dn * fabs(rp));       // unscaled: E^-1 cca = D
          v[2] = max_bits(v[2], fabs(Dn * z[i]));
          v[3] = max_bits(v[3], fabs(Dn * Ah[i]));
          v[4] = max_bits(v[4], fabs(cca[i] * z[i]));
          v[5] = max_bits(v[5], fabs(cca[i] * Ah[i]));
        }
        // P_ x = c D (R2 D x + G' w) ; A_' y
        const double* wv = &vo[6 * k];
        const double f0 = lc[0] * u[0], f1 = lc[1] * u[1], f2 = lc[2] * u[2], f3 = lc[3] * u[3], f4 = lc[4] * u[4];
        const double qb[3] = {lc[11], lc[12], lc[13]};
        const double Aty[3] = {f0 + f1, f2 + f3, f4 + (tzx * (f0 - f1) + tzy * (f2 - f3))};
        const double r2v[3] = {bp.Rd[3 * lg], bp.Rd[3 * lg + 1], bp.Rd[3 * lg + 2]};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const double gtw = top[0][q] * wv[0] + top[1][q] * wv[1] + top[2][q] * wv[2] + inv_m * wv[3 + q];
          const double Px = cs * D[q] * (r2v[q] * (D[q] * x[q]) + gtw);
          const double Dinv = 1.0 / D[q];
          const double rd = Px + qb[q] + Aty[q];
          v[6] = max_bits(v[6], fabs(rd));
          v[7] = max_bits(v[7], fabs(Dinv * rd));
          v[8] = max_bits(v[8], max_bits(max_bits(fabs(Dinv * qb[q]), fabs(Dinv * Aty[q])), fabs(Dinv * Px)));
          v[9] = max_bits(v[9], max_bits(max_bits(fabs(qb[q]), fabs(Aty[q])), fabs(Px)));
        }
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] = max_bits(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
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
          for (int w = 1; w < kWrcWarps; ++w) t = max_bits(t, sm.red[w * 16 + i]);
          mres[i] = t;
        }
        const double pri = mres[1], dua = cinv * mres[7];
        const double eps_pri = sp.eps_abs + sp.eps_rel * max_bits(mres[2], mres[3]);
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
      WRP(7);  // residual check
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
          const double rn = rho_row(i, rho), ro = rho_row(i, rho_cur);
          if (isleg) {
            const double cc_ = gcca[i];
            zu[5 + i] *= ro / rn;   // y stays, uh = y / (rho cca) follows the new rho
            lc[i] = rn * cc_ * cc_;
          }
        }
        rho_cur = rho;
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
      for (int i = 0; i < 5; ++i) fin = fin && isfinite(zu[i]) && isfinite(zu[5 + i]);
      const int all_ok = __syncthreads_and(fin || !isleg);
      if (isleg) {
#pragma unroll
        for (int q = 0; q < 3; ++q) ws[kWX + j0 + q] = x[q];   // (q went into the slot when it was formed)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const double cc_ = gcca[i];
          ws[kWZ + r0 + i] = cc_ * zu[i];            // OSQP's scaled z, y
          ws[kWY + r0 + i] = rho_row(i, rho_cur) * cc_ * zu[5 + i];
        }
      }
      if (tid == 0) {
        ws[kWRho] = sm.scal[2];
        ws[kWLive] = (all_ok && isfinite(sm.scal[2])) ? 1.0 : 0.0;
      }
    }
    const double f0 = sm.Dp[j0] * x[0], f1 = sm.Dp[j0 + 1] * x[1], f2 = sm.Dp[j0 + 2] * x[2];
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
#ifdef WRC_PROF
    if (tid == 0 && p < 2)
      printf("WRCPROF p %d iters %d: ruiz %lld factor: local %lld riccati %lld phi %lld | u %lld bwd %lld eb %lld fwd %lld xrow %lld check %lld\n", p, iter,
             pc_[0], pc_[8], pc_[9], pc_[1], pc_[2], pc_[3], pc_[4], pc_[5], pc_[6], pc_[7]);
#endif
    if (tid == 0) {
      results[p].status = status;
      results[p].iters = iter;
      results[p].rho_updates = rho_updates;
      results[p].pri_res = (float)pri_res_out;
    }
  }
}

}  // namespace mpcb200
