// Structured ADMM solve of the condensed MPC QP (any horizon H): the same OSQP iteration as
// admm_kernel.cuh (Ruiz equilibration, per-row rho, K x~ = sigma x - q + A'(rho z - y), relaxation,
// residual termination every check_termination iterations, rho adaptation with refactorisation;
// OSQP 0.6.x as driven by A1RobotControl.cpp:522-561), but the linear system is never formed:
//
//   K = c D (R + B_qp' Q B_qp) D + sigma I + A' diag(rho) A
//
// is the Hessian of a finite-horizon LQR problem (states X_k+1 = A_d X_k + Bs_k u_k from X_0 = 0,
// Bs_k = B_d(k) diag(D_k), stage cost c Q on X_k+1, input cost Dk = c D R D + sigma I + G_k), so
// K x~ = r is solved EXACTLY by a Riccati recursion:
//   factor (once per rho):  Pi_H = cQ;  M_k = Dk + Bs' Pi_k+1 Bs,  K_k = M_k^-1 Bs' Pi_k+1 A,
//                           L_k = A - Bs K_k,  Pi_k = cQ + A' Pi_k+1 L_k                 O(H 13^3)
//   solve (per iteration):  p_k = L_k' p_k+1 + K_k' r_k  (backward),  g_k = -M_k^-1 (Bs' p_k+1 - r_k),
//                           X_k+1 = L_k X_k + Bs g_k (forward),  u_k = -K_k X_k + g_k     O(H 13^2)
// against n^3 / n^2 for the dense inverse.  For H = 30 that is 1/40 of the flops and, decisive on
// this machine, 110 KB of per-step gains in shared memory instead of a 1.04 MB inverse streamed
// from L2 every iteration.  P x for the residuals is the same structure without feedback.  The
// explicit Hessian (from the build kernel, in HBM) is only read by the Ruiz passes.
//
// One CTA of 256 threads per problem (problems pulled from an atomic counter).  The recursions are
// the serial part: steps are grouped by kRicGroup = 5 and the factorisation also forms the group
// transitions Phi_j = L_(5j+4) ... L_(5j), so that each recursion runs as three sweeps -- all groups at
// once from a zero boundary (one group per half warp), the true group boundaries on one warp through
// Phi_j, the groups again from their true boundary -- 4 + 5 + 4 dependent steps at H = 30 instead of
// 30 (ric_chain).  Everything else is one- or two-outputs-per-thread phases over the horizon.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_kernels.cuh"

namespace mpcb200 {

constexpr int kRicThreads = 256;
constexpr int kRicWarps = kRicThreads / 32;
constexpr int kVS = 14;  // row stride of the recursion vectors in shared memory
constexpr int kRicGroup = 5;  // steps per recursion group (H must be a multiple)

template <int H>
struct RicSmem {
  static constexpr int n = 12 * H, m = 20 * H;
  double A[169];
  double A5[169];  // A_d^5: group transition of the open-loop recursions of the residual check
  alignas(16) double Bs[H][156];   // scaled input matrices Bs_k = B_d(k) diag(D_k), 13 x 12 row-major
  alignas(16) double Kk[H][156];   // gains K_k (12 x 13) stored TRANSPOSED, 13 rows of 12: Kk[i * 12 + a] = K[a][i]
  double Lk[H][169];   // closed loop A - Bs K, 13 x 13
  alignas(16) double Mi[H][144];   // M_k^-1, 12 x 12
  double Pi[169];
  // factor scratch W1|W2|W3|W4|Mm (794 doubles) while the recursion is being factored; afterwards the
  // group transition matrices Phi_j = L_(5j+4) ... L_(5j), 13 x 13 each, used by every iteration
  static constexpr int kFs = (H / 5) * 169 > 794 ? (H / 5) * 169 : 794;
  double fs[kFs];
  alignas(16) double x[n];
  alignas(16) double xt[n];
  alignas(16) double rhs[n];
  double qb[n], Dv[n], Px[n];  // xt doubles as the new-D scratch of the Ruiz passes
  double z[m], y[m], rv[m], cca[m], ccz[m], Ev[m];
  float lb[m], ub[m];  // UNSCALED bounds exactly as given; scaled by E (f64) where they are used
  alignas(16) double pv[(H + 2) * kVS];  // costates, one row of kVS = 14 doubles per step (16 B aligned)
  alignas(16) double Xv[(H + 1) * kVS];  // states
  double tv[H * 13];
  alignas(16) double gv[H * 12];
  alignas(16) double wv[H * 12];
  double red[16 * kRicWarps];
  double scal[8];  // 0:c 1:cinv 2:rho 4:pri_res
  int flags[8];    // 0:done 1:status 2:refactor 3:problem index
};

__device__ __forceinline__ double ric_limit_scaling(double v) {
  v = v < 1e-4 ? 1.0 : v;
  return v > 1e4 ? 1e4 : v;
}

// max of NON-NEGATIVE doubles through their bit patterns (ordered like the values); bit-identical
__device__ __forceinline__ double ric_max_nn(double a, double b) {
  return (__double_as_longlong(a) > __double_as_longlong(b)) ? a : b;
}

// block-wide max or sum of one value per thread; result to every thread
template <bool kMax>
__device__ __forceinline__ double ric_block_reduce(double v, double* red, int slot) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, v, o);
    v = kMax ? fmax(v, t) : v + t;
  }
  if ((threadIdx.x & 31) == 0) red[slot * kRicWarps + (threadIdx.x >> 5)] = v;
  __syncthreads();
  double r[kRicWarps];
#pragma unroll
  for (int w = 0; w < kRicWarps; ++w) r[w] = red[slot * kRicWarps + w];
#pragma unroll
  for (int st = 1; st < kRicWarps; st *= 2)
#pragma unroll
    for (int w = 0; w + st < kRicWarps; w += 2 * st) r[w] = kMax ? fmax(r[w], r[w + st]) : r[w] + r[w + st];
  return r[0];
}

// dot product of two 16-byte aligned runs of 12 doubles: six 16-byte loads each, two accumulators.
// Matrix rows are 96 bytes apart, so rows r and r + 4 start in the same bank group; `rot` (0 or 1,
// bit 2 of the linear row index) makes the second of each such pair walk its chunks one position
// ahead, which keeps the eight lanes of a quarter warp on eight different 16-byte bank groups.
__device__ __forceinline__ double ric_dot12(const double* a, const double* b, int rot) {
  const double2* a2 = reinterpret_cast<const double2*>(a);
  const double2* b2 = reinterpret_cast<const double2*>(b);
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int h = 0; h < 6; ++h) {
    const int c = (h == 5) ? (rot ? 0 : 5) : h + rot;
    const double2 x = a2[c], y = b2[c];
    s0 = fma(x.x, y.x, s0);
    s1 = fma(x.y, y.y, s1);
  }
  return s0 + s1;
}

// two adjacent outputs  (o0, o1) = sum_i M[i * 12 + a .. a + 1] v[i],  i < 13:  M 13 x 12 row-major with a
// even (16-byte loads of the pair), v a 16-byte aligned row of kVS = 14 doubles
__device__ __forceinline__ void ric_pair13(const double* Mcol, const double* vrow, double& o0, double& o1) {
  const double2* v2 = reinterpret_cast<const double2*>(vrow);
  double v[14];
#pragma unroll
  for (int h = 0; h < 7; ++h) { const double2 t = v2[h]; v[2 * h] = t.x; v[2 * h + 1] = t.y; }
  double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    const double2 ma = *reinterpret_cast<const double2*>(Mcol + i * 12);
    const double2 mb = *reinterpret_cast<const double2*>(Mcol + (i + 1) * 12);
    s0 = fma(ma.x, v[i], s0);
    s1 = fma(ma.y, v[i], s1);
    t0 = fma(mb.x, v[i + 1], t0);
    t1 = fma(mb.y, v[i + 1], t1);
  }
  const double2 mc = *reinterpret_cast<const double2*>(Mcol + 12 * 12);
  s0 = fma(mc.x, v[12], s0);
  s1 = fma(mc.y, v[12], s1);
  o0 = s0 + t0;
  o1 = s1 + t1;
}

__device__ __forceinline__ double ric_lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 ric_lds2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// One 13-lane recursion  out_s = M_s in_s + add_s,  s = 0 .. nsteps-1, on a half warp.  Every pointer
// is per lane:  Mp + s m_step + j SJ  is element j of this lane's row (SJ = 1) or column (SJ = 13, for
// M') of M_s;  vin + s v_step  the 13-vector in_s (a 16-byte aligned row of kVS doubles; out_(s-1) for
// s > 0);  vout + s o_step  this lane's element of out_s;  addp + s a_step  its addend.
// A step issues, IN THIS ORDER (the loads are volatile asm so that the order survives scheduling):
// the seven 16-byte loads of the vector, which are the critical path behind the previous step's
// store; the matrix elements and addend of the NEXT step into the other register set (always -- one
// step past the end reads shared memory that is allocated and ignores it, which keeps the step free
// of branches); 13 FMAs on four accumulators, the addend seeding one; the store; the warp barrier.
// One shared body per direction keeps the iteration loop inside the instruction cache.
template <int SJ>
__device__ __noinline__ void ric_chain(const double* Mp, int m_step, const double* vin, int v_step,
                                       double* vout, int o_step, const double* addp, int a_step,
                                       int nsteps, bool wr) {
  uint32_t m_a = (uint32_t)__cvta_generic_to_shared(Mp);
  uint32_t v_a = (uint32_t)__cvta_generic_to_shared(vin);
  uint32_t d_a = (uint32_t)__cvta_generic_to_shared(addp);
  const int m_b = m_step * 8, v_b = v_step * 8, d_b = a_step * 8;
  auto step = [&](const double (&Lr)[13], double ad, double (&Ln)[13], double& adn) {
    double v[14];
#pragma unroll
    for (int h = 0; h < 7; ++h) { const double2 t = ric_lds2(v_a + 16 * h); v[2 * h] = t.x; v[2 * h + 1] = t.y; }
    v_a += v_b;
    m_a += m_b;
    d_a += d_b;
#pragma unroll
    for (int j = 0; j < 13; ++j) Ln[j] = ric_lds(m_a + 8 * SJ * j);
    adn = ric_lds(d_a);
    double s0 = ad, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int j = 0; j < 12; j += 4) {
      s0 = fma(Lr[j], v[j], s0);
      s1 = fma(Lr[j + 1], v[j + 1], s1);
      s2 = fma(Lr[j + 2], v[j + 2], s2);
      s3 = fma(Lr[j + 3], v[j + 3], s3);
    }
    s1 = fma(Lr[12], v[12], s1);
    if (wr) *vout = (s0 + s1) + (s2 + s3);
    vout += o_step;
    __syncwarp();
  };
  double LA[13], LB[13], aA, aB;
#pragma unroll
  for (int j = 0; j < 13; ++j) LA[j] = ric_lds(m_a + 8 * SJ * j);
  aA = ric_lds(d_a);
  __syncwarp();
  int left = nsteps;
#pragma unroll 1
  for (; left >= 2; left -= 2) {
    step(LA, aA, LB, aB);
    step(LB, aB, LA, aA);
  }
  if (left == 1) step(LA, aA, LB, aB);
}

template <int H>
__global__ void __launch_bounds__(kRicThreads)
riccati_solve_kernel(const double* __restrict__ P_all, size_t p_stride, int p_row_stride,
                     const double* __restrict__ q_all, const float* __restrict__ l_all,
                     const float* __restrict__ u_all, const double* __restrict__ model_all,
                     const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                     float* __restrict__ x_all, int num, int* __restrict__ counter,
                     double* __restrict__ warm, int warm_stride,
                     const __grid_constant__ BuildParams bp, const __grid_constant__ SolveParams sp) {
  constexpr int n = 12 * H, m = 20 * H;
  // warm-start slot of one robot (the solver the controller keeps alive, A1RobotControl.cpp:522-538):
  // scaled x | previous unscaled q | scaled z | scaled y | rho | live -- the layout of admm_kernel.cuh
  constexpr int kWX = 0, kWQ = n, kWZ = 2 * n, kWY = 2 * n + m, kWRho = 2 * n + 2 * m, kWLive = kWRho + 1;
  extern __shared__ __align__(16) unsigned char ric_smem_raw[];
  RicSmem<H>& sm = *reinterpret_cast<RicSmem<H>*>(ric_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double sigma = sp.sigma, alpha = sp.alpha, mu = sp.mu;
  double* const W1 = sm.fs;
  double* const W2 = sm.fs + 169;
  double* const W3 = sm.fs + 338;
  double* const W4 = sm.fs + 494;
  double* const Mm = sm.fs + 650;
  double (*const Phi)[169] = reinterpret_cast<double (*)[169]>(sm.fs);
  constexpr int kG = H / kRicGroup;  // groups of kRicGroup consecutive steps
  // recursion slots: two per warp (lanes 0-12 and 16-28); slot q works on group q
  const int hl = lane & 15, slot = 2 * warp + (lane >> 4);
  const bool slot_on = hl < 13 && slot < kG;
  const int sl = hl < 13 ? hl : 0, sg = slot < kG ? slot : 0;

  // Profiling build only (-DRIC_PROF, or `#define RIC_PROF 1` above; run scripts/prof_ric.py): thread 0
  // accumulates clock64 per phase and prints the totals of the first two problems.  Compiled out
  // otherwise: RP(i) is empty.
#ifdef RIC_PROF
  long long pc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pm_ = 0;
#define RP(i) do { if (tid == 0) { const long long t_ = clock64(); pc_[i] += t_ - pm_; pm_ = t_; } } while (0)
#else
#define RP(i) do {} while (0)
#endif
  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;
#ifdef RIC_PROF
    if (tid == 0) pm_ = clock64();
#endif
    const double* Pg = P_all + size_t(p) * p_stride;
    const double* model = model_all + size_t(p) * (169 + H * 156);

    // a live slot makes this an update + warm solve (osqp_update_P / _lin_cost / _bounds, then solve):
    // the equilibration still sees the PREVIOUS gradient, x, z, y and rho carry over
    double* const ws = warm ? warm + size_t(p) * warm_stride : nullptr;
    const bool live = warm && ws[kWLive] != 0.0;
    const double rho0 = live ? ws[kWRho] : sp.rho;
    // ---- load: A_d, bounds, gradient ----
    for (int i = tid; i < 169; i += kRicThreads) sm.A[i] = model[i];
    for (int i = tid; i < n; i += kRicThreads) {
      sm.Dv[i] = 1.0;
      sm.x[i] = live ? ws[kWX + i] : 0.0;
    }
    for (int i = tid; i < m; i += kRicThreads) {
      sm.Ev[i] = 1.0;
      sm.z[i] = live ? ws[kWZ + i] : 0.0;
      sm.y[i] = live ? ws[kWY + i] : 0.0;
    }
    if (tid == 0) {
      sm.scal[0] = 1.0;
      sm.scal[2] = rho0;
      sm.flags[0] = 0;
      sm.flags[1] = MPC_STATUS_UNSOLVED;
    }
    __syncthreads();
    {  // A_d^kRicGroup by repeated products through the (still unused) factor scratch
      const double* src = sm.A;
#pragma unroll 1
      for (int s_ = 1; s_ < kRicGroup; ++s_) {
        double* dst = (s_ == kRicGroup - 1) ? sm.A5 : ((s_ & 1) ? W1 : W2);
        if (tid < 169) {
          const int r = tid / 13, cc = tid - 13 * r;
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 13; ++i) acc = fma(sm.A[r * 13 + i], src[i * 13 + cc], acc);
          dst[tid] = acc;
        }
        __syncthreads();
        src = dst;
      }
    }

    // ---- modified Ruiz equilibration (osqp scaling.c scale_data), P read from global ----
    double c_run = 1.0;
    // row norms of the scaled Hessian, max_j |P_rj| D_j (its column norms, by symmetry): warp per row
    // Only the upper triangle is read: element (r, j), j > r, serves row r (times D_j) and, by symmetry,
    // row j (times D_r), so a pass moves half the bytes and the 148 resident Hessians (74 MB of upper
    // triangles) stay in L2 between passes.  Row maxima by shuffle, column maxima in registers per
    // warp; both are merged in shared memory as integer maxima of the (non-negative) bit patterns.
    auto p_row_norms = [&](bool accumulate, double& psum, double& qmax) {
      constexpr int kNJ = (n + 31) / 32;
      unsigned long long* pxb = reinterpret_cast<unsigned long long*>(sm.Px);
      for (int j = tid; j < n; j += kRicThreads) pxb[j] = 0ull;
      __syncthreads();
      double cm[kNJ];
#pragma unroll
      for (int t = 0; t < kNJ; ++t) cm[t] = 0.0;
      // rows rp and n - 1 - rp together hold n + 1 upper-triangle elements: every warp iteration does
      // the same work.  The loads of the NEXT pair are in flight while this one is reduced (two
      // register sets, loop unrolled by two): a pass is bound by the latency of one pair otherwise.
      auto load_pair = [&](int rp, double (&ea)[kNJ], double (&eb)[kNJ]) {
        const bool on = rp < n / 2;
        const int ra = on ? rp : 0, rb = n - 1 - ra;
        const double* rowa = Pg + size_t(ra) * p_row_stride;
        const double* rowb = Pg + size_t(rb) * p_row_stride;
#pragma unroll
        for (int t = 0; t < kNJ; ++t) {
          const int j = lane + 32 * t;
          ea[t] = (on && j >= ra && j < n) ? rowa[j] : 0.0;
          eb[t] = (on && j >= rb && j < n) ? rowb[j] : 0.0;
        }
      };
      auto reduce_pair = [&](int rp, const double (&ea)[kNJ], const double (&eb)[kNJ]) {
        if (rp >= n / 2) return;
        const int ra = rp, rb = n - 1 - rp;
        const double da = sm.Dv[ra], db = sm.Dv[rb];
        double ma = 0.0, mb = 0.0;
#pragma unroll
        for (int t = 0; t < kNJ; ++t) {
          const int j = lane + 32 * t;
          const double d = (j < n) ? sm.Dv[j] : 0.0;
          const double fa = fabs(ea[t]), fb = fabs(eb[t]);
          // maxima of non-negative doubles through their bit patterns (integer compares, not DSETP)
          ma = ric_max_nn(ma, fa * d);
          mb = ric_max_nn(mb, fb * d);
          // the diagonal element belongs to its row only
          cm[t] = ric_max_nn(cm[t], ric_max_nn((j > ra) ? fa * da : 0.0, (j > rb) ? fb * db : 0.0));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ma = ric_max_nn(ma, __shfl_xor_sync(0xffffffffu, ma, o));
          mb = ric_max_nn(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        }
        if (lane == 0) atomicMax(&pxb[ra], (unsigned long long)__double_as_longlong(ma));
        if (lane == 1) atomicMax(&pxb[rb], (unsigned long long)__double_as_longlong(mb));
      };
      {
        double eA[kNJ], fA[kNJ], eB[kNJ], fB[kNJ];
        load_pair(warp, eA, fA);
#pragma unroll 1
        for (int rp = warp; rp < n / 2; rp += 2 * kRicWarps) {
          load_pair(rp + kRicWarps, eB, fB);
          reduce_pair(rp, eA, fA);
          load_pair(rp + 2 * kRicWarps, eA, fA);
          reduce_pair(rp + kRicWarps, eB, fB);
        }
      }
#pragma unroll
      for (int t = 0; t < kNJ; ++t) {
        const int j = lane + 32 * t;
        if (j < n && cm[t] > 0.0) atomicMax(&pxb[j], (unsigned long long)__double_as_longlong(cm[t]));
      }
      __syncthreads();
      for (int r = tid; r < n; r += kRicThreads) {
        const double v = c_run * sm.Dv[r] * sm.Px[r];
        sm.Px[r] = v;
        if (accumulate) {
          psum += v;
          qmax = fmax(qmax, fabs(c_run * sm.Dv[r] * (live ? ws[kWQ + r] : q_all[size_t(p) * n + r])));
        }
      }
    };
    if (sp.scaling > 0) {
      double dummy_s = 0.0, dummy_q = 0.0;
      p_row_norms(false, dummy_s, dummy_q);  // c = 1, D = 1
      __syncthreads();
      for (int it = 0; it < sp.scaling; ++it) {
        // new D from the column norms of [P; A] (old E), new E from the row norms of A (old D)
        for (int j = tid; j < n; j += kRicThreads) {
          const int ls = j / 3, vc = j - 3 * ls;
          const double* Er = &sm.Ev[5 * ls];
          double nA;
          if (vc == 0) nA = fmax(Er[0], Er[1]);
          else if (vc == 1) nA = fmax(Er[2], Er[3]);
          else nA = fmax(mu * fmax(fmax(Er[0], Er[1]), fmax(Er[2], Er[3])), Er[4]);
          nA *= sm.Dv[j];
          sm.xt[j] = sm.Dv[j] * rsqrt(ric_limit_scaling(fmax(sm.Px[j], nA)));
        }
        __syncthreads();
        for (int i = tid; i < m; i += kRicThreads) {
          const int ls = i / 5, pos = i - 5 * ls;
          const double Dx = sm.Dv[3 * ls], Dy = sm.Dv[3 * ls + 1], Dz = sm.Dv[3 * ls + 2];
          const double nrow = (pos == 4) ? Dz : fmax((pos < 2) ? Dx : Dy, mu * Dz);
          const double E = sm.Ev[i];
          sm.Ev[i] = E * rsqrt(ric_limit_scaling(E * nrow));
        }
        __syncthreads();
        for (int j = tid; j < n; j += kRicThreads) sm.Dv[j] = sm.xt[j];
        __syncthreads();
        // cost normalisation with the new D and the old c
        double psum = 0.0, qmax = 0.0;
        p_row_norms(true, psum, qmax);
        const double s_all = ric_block_reduce<false>(psum, sm.red, 0);
        const double q_all_max = ric_block_reduce<true>(qmax, sm.red, 1);
        const double ct = 1.0 / ric_limit_scaling(fmax(s_all / (double)n, ric_limit_scaling(q_all_max)));
        c_run *= ct;
        for (int j = tid; j < n; j += kRicThreads) sm.Px[j] *= ct;  // norms under the new c
        __syncthreads();
      }
    }
    RP(0);  // load + Ruiz
    const double c = c_run, cinv = 1.0 / c_run;
    // ---- scaled data ----
    for (int j = tid; j < n; j += kRicThreads) sm.qb[j] = c * sm.Dv[j] * q_all[size_t(p) * n + j];
    for (int i = tid; i < m; i += kRicThreads) {
      const int ls = i / 5, pos = i - 5 * ls;
      const double E = sm.Ev[i];
      const double lo = (double)l_all[size_t(p) * m + i] * E, hi = (double)u_all[size_t(p) * m + i] * E;
      sm.lb[i] = l_all[size_t(p) * m + i];
      sm.ub[i] = u_all[size_t(p) * m + i];
      const double Dlat = sm.Dv[3 * ls + ((pos < 2) ? 0 : 1)], Dz = sm.Dv[3 * ls + 2];
      sm.cca[i] = (pos == 4) ? 0.0 : E * Dlat;
      sm.ccz[i] = (pos == 4) ? E * Dz : ((pos & 1) ? -mu : mu) * E * Dz;
      int ctype = 0;
      if (lo < -MPC_INFTY * 1e-4 && hi > MPC_INFTY * 1e-4) ctype = -1;
      else if (hi - lo < 1e-4) ctype = 1;
      const double rv = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho0 : rho0;
      sm.rv[i] = rv;
      // the constraint type is re-derived from the bounds at every rho update (same test)
    }
    for (int idx = tid; idx < H * 156; idx += kRicThreads) {
      const int k = idx / 156, e = idx - 156 * k, a = e % 12;
      sm.Bs[k][e] = model[169 + idx] * sm.Dv[12 * k + a];
    }
    __syncthreads();
    // first rhs = sigma x - q + A'(rho z - y) (= -q on a cold start: x = z = y = 0)
    for (int j = tid; j < n; j += kRicThreads) {
      const int ls = j / 3, vc = j - 3 * ls;
      const int r0 = 5 * ls;
      double s_ = 0.0;
      if (live) {
        if (vc == 0) {
          s_ = sm.cca[r0] * (sm.rv[r0] * sm.z[r0] - sm.y[r0]) + sm.cca[r0 + 1] * (sm.rv[r0 + 1] * sm.z[r0 + 1] - sm.y[r0 + 1]);
        } else if (vc == 1) {
          s_ = sm.cca[r0 + 2] * (sm.rv[r0 + 2] * sm.z[r0 + 2] - sm.y[r0 + 2]) +
               sm.cca[r0 + 3] * (sm.rv[r0 + 3] * sm.z[r0 + 3] - sm.y[r0 + 3]);
        } else {
#pragma unroll
          for (int rw = 0; rw < 5; ++rw) s_ = fma(sm.ccz[r0 + rw], sm.rv[r0 + rw] * sm.z[r0 + rw] - sm.y[r0 + rw], s_);
        }
      }
      sm.rhs[j] = live ? (sigma * sm.x[j] - sm.qb[j] + s_) : -sm.qb[j];
    }
    __syncthreads();

    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    bool need_factor = true;
    double rho_cur = rho0, rinv_in = 1.0 / rho0, rinv_eq = 1.0 / (1e3 * rho0);
    const double rinv_free = 1.0 / 1e-6;
    for (iter = 1; iter <= sp.max_iter; ++iter) {
      if (need_factor) {
        need_factor = false;
        // ---- Riccati factorisation ----
        for (int i = tid; i < 169; i += kRicThreads) sm.Pi[i] = (i / 13 == i % 13) ? c * bp.Qd[i / 13] : 0.0;
        __syncthreads();
        for (int k = H - 1; k >= 0; --k) {
          const double* Bk = sm.Bs[k];
          // W3 = Bs' Pi (12 x 13)
          for (int idx = tid; idx < 156; idx += kRicThreads) {
            const int a = idx / 13, j = idx - 13 * a;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 13; ++i) s = fma(Bk[i * 12 + a], sm.Pi[i * 13 + j], s);
            W3[idx] = s;
          }
          __syncthreads();
          // Mm = Dk + W3 Bs (12 x 12) ; W4 = W3 A (12 x 13)
          for (int idx = tid; idx < 144 + 156; idx += kRicThreads) {
            if (idx < 144) {
              const int a = idx / 12, b = idx - 12 * a;
              double s = 0.0;
#pragma unroll
              for (int i = 0; i < 13; ++i) s = fma(W3[a * 13 + i], Bk[i * 12 + b], s);
              // input cost block: c D R D + sigma on the diagonal, G = A' rho A per leg (3 x 3)
              if (a / 3 == b / 3) {
                const int ls = 4 * k + a / 3, ra = a % 3, rb = b % 3;
                const double* r_ = &sm.rv[5 * ls];
                const double* ca = &sm.cca[5 * ls];
                const double* cz = &sm.ccz[5 * ls];
                double gsum = 0.0;
                // rows 0,1 touch (x, z); rows 2,3 touch (y, z); row 4 touches z
#pragma unroll
                for (int rw = 0; rw < 5; ++rw) {
                  const double fa = (rw < 2) ? ((ra == 0) ? ca[rw] : (ra == 2) ? cz[rw] : 0.0)
                                   : (rw < 4) ? ((ra == 1) ? ca[rw] : (ra == 2) ? cz[rw] : 0.0)
                                              : ((ra == 2) ? cz[rw] : 0.0);
                  const double fb = (rw < 2) ? ((rb == 0) ? ca[rw] : (rb == 2) ? cz[rw] : 0.0)
                                   : (rw < 4) ? ((rb == 1) ? ca[rw] : (rb == 2) ? cz[rw] : 0.0)
                                              : ((rb == 2) ? cz[rw] : 0.0);
                  gsum = fma(r_[rw] * fa, fb, gsum);
                }
                s += gsum;
                if (a == b) {
                  const double d = sm.Dv[12 * k + a];
                  s += c * d * bp.Rd[a] * d + sigma;
                }
              }
              Mm[idx] = s;
            } else {
              const int e = idx - 144, a = e / 13, j = e - 13 * a;
              double s = 0.0;
#pragma unroll
              for (int i = 0; i < 13; ++i) s = fma(W3[a * 13 + i], sm.A[i * 13 + j], s);
              W4[e] = s;
            }
          }
          __syncthreads();
          // Mi = Mm^-1: Gauss-Jordan sweep on the SPD 12 x 12 block, warp 0, lane r holds row r in
          // registers; the pivot row travels by shuffle (no shared-memory round trips between pivots)
          if (warp == 0) {
            const int r = lane < 12 ? lane : 0;
            double row[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) row[j] = Mm[r * 12 + j];
#pragma unroll
            for (int pv_ = 0; pv_ < 12; ++pv_) {
              double rowp[12];
#pragma unroll
              for (int j = 0; j < 12; ++j) rowp[j] = __shfl_sync(0xffffffffu, row[j], pv_);
              const double d = 1.0 / rowp[pv_];
              const bool piv = (lane == pv_);
              const double f = row[pv_] * d;
#pragma unroll
              for (int j = 0; j < 12; ++j) {
                const double upd = (j == pv_) ? -f : row[j] - f * rowp[j];
                const double prw = (j == pv_) ? d : rowp[j] * d;
                row[j] = piv ? prw : upd;
              }
            }
            if (lane < 12) {
#pragma unroll
              for (int j = 0; j < 12; ++j) sm.Mi[k][lane * 12 + j] = row[j];
            }
          }
          __syncthreads();
          // Kk = Mi W4 (12 x 13)
          for (int idx = tid; idx < 156; idx += kRicThreads) {
            const int a = idx / 13, j = idx - 13 * a;
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 12; ++b) s = fma(sm.Mi[k][a * 12 + b], W4[b * 13 + j], s);
            sm.Kk[k][j * 12 + a] = s;
          }
          __syncthreads();
          // Lk = A - Bs Kk (13 x 13)
          for (int idx = tid; idx < 169; idx += kRicThreads) {
            const int i = idx / 13, j = idx - 13 * i;
            sm.Lk[k][idx] = sm.A[idx] - ric_dot12(&Bk[i * 12], &sm.Kk[k][j * 12], 0);
          }
          __syncthreads();
          // W1 = Pi Lk ; then Pi <- sym(cQ + A' W1)
          for (int idx = tid; idx < 169; idx += kRicThreads) {
            const int i = idx / 13, j = idx - 13 * i;
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < 13; ++t) s = fma(sm.Pi[i * 13 + t], sm.Lk[k][t * 13 + j], s);
            W1[idx] = s;
          }
          __syncthreads();
          for (int idx = tid; idx < 169; idx += kRicThreads) {
            const int i = idx / 13, j = idx - 13 * i;
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < 13; ++t) s = fma(sm.A[t * 13 + i], W1[t * 13 + j], s);
            W2[idx] = s;
          }
          __syncthreads();
          for (int idx = tid; idx < 169; idx += kRicThreads) {
            const int i = idx / 13, j = idx - 13 * i;
            sm.Pi[idx] = 0.5 * (W2[idx] + W2[j * 13 + i]) + ((i == j) ? c * bp.Qd[i] : 0.0);
          }
          __syncthreads();
        }
        // group transitions Phi_j = L_(gj+g-1) ... L_(gj): one thread per column, the column in registers
        if (tid < kG * 13) {
          const int gj = tid / 13, col = tid - 13 * gj;
          double T[13], Tn[13];
#pragma unroll
          for (int i = 0; i < 13; ++i) T[i] = sm.Lk[kRicGroup * gj][i * 13 + col];
#pragma unroll 1
          for (int s_ = 1; s_ < kRicGroup; ++s_) {
            const double* Ls = sm.Lk[kRicGroup * gj + s_];
#pragma unroll
            for (int r = 0; r < 13; ++r) {
              double acc = 0.0;
#pragma unroll
              for (int i = 0; i < 13; ++i) acc = fma(Ls[r * 13 + i], T[i], acc);
              Tn[r] = acc;
            }
#pragma unroll
            for (int i = 0; i < 13; ++i) T[i] = Tn[i];
          }
#pragma unroll
          for (int r = 0; r < 13; ++r) Phi[gj][r * 13 + col] = T[r];
        }
        __syncthreads();
      }
      RP(1);  // (factor when it ran, else loop overhead)
      // ---- x~ = K^-1 rhs by the two recursions ----
      for (int idx = tid; idx < H * 13; idx += kRicThreads) {  // t_k = K_k' r_k
        const int k = idx / 13, i = idx - 13 * k;
        sm.tv[idx] = ric_dot12(&sm.Kk[k][i * 12], &sm.rhs[12 * k], (idx >> 2) & 1);
      }
      if (tid < 13) sm.pv[H * kVS + tid] = 0.0;
      __syncthreads();
      RP(2);  // t phase
      {
        // p_k = L_k' p_k+1 + t_k (p_H = 0) in three sweeps of 4 + (kG - 1) + 4 dependent steps instead of
        // H: (1) every group runs its own recursion from a zero terminal costate, all groups at once on
        // the half-warp slots; (2) warp 0 carries the group boundaries, p_gj = Phi_j' p_g(j+1) + (1);
        // (3) the groups redo their interior steps from the true boundary value.
        constexpr int g = kRicGroup;
        const int kl = g * sg + g - 1;  // last step of this slot's group
        if (slot_on) sm.pv[kl * kVS + sl] = sm.tv[kl * 13 + sl];
        if (2 * warp < kG)
          ric_chain<13>(&sm.Lk[kl - 1][sl], -169, &sm.pv[kl * kVS], -kVS, &sm.pv[(kl - 1) * kVS + sl], -kVS,
                    &sm.tv[(kl - 1) * 13 + sl], -13, g - 1, slot_on);
        __syncthreads();
        if (warp == 0 && kG > 1) {
          const int l0 = lane < 13 ? lane : 0;
          ric_chain<13>(&Phi[kG - 2][l0], -169, &sm.pv[g * (kG - 1) * kVS], -g * kVS,
                    &sm.pv[g * (kG - 2) * kVS + l0], -g * kVS, &sm.pv[g * (kG - 2) * kVS + l0], -g * kVS,
                    kG - 1, lane < 13);
        }
        __syncthreads();
        if (2 * warp < kG)
          ric_chain<13>(&sm.Lk[kl][sl], -169, &sm.pv[(kl + 1) * kVS], -kVS, &sm.pv[kl * kVS + sl], -kVS,
                    &sm.tv[kl * 13 + sl], -13, g - 1, slot_on);
      }
      RP(3);  // backward chain
      __syncthreads();
      for (int idx = tid; idx < H * 6; idx += kRicThreads) {  // w_k = Bs' p_k+1 - r_k, two outputs per thread
        const int k = idx / 6, a = 2 * (idx - 6 * k);
        double o0, o1;
        ric_pair13(&sm.Bs[k][a], &sm.pv[(k + 1) * kVS], o0, o1);
        const double2 r2 = *reinterpret_cast<const double2*>(&sm.rhs[12 * k + a]);
        *reinterpret_cast<double2*>(&sm.wv[12 * k + a]) = make_double2(o0 - r2.x, o1 - r2.y);
      }
      __syncthreads();
      for (int idx = tid; idx < H * 12; idx += kRicThreads) {  // g_k = -M_k^-1 w_k
        const int k = idx / 12, a = idx - 12 * k;
        sm.gv[idx] = -ric_dot12(&sm.Mi[k][a * 12], &sm.wv[12 * k], (idx >> 2) & 1);
      }
      __syncthreads();
      for (int idx = tid; idx < H * 13; idx += kRicThreads) {  // b_k = Bs g_k (into tv)
        const int k = idx / 13, i = idx - 13 * k;
        sm.tv[idx] = ric_dot12(&sm.Bs[k][i * 12], &sm.gv[12 * k], (idx >> 2) & 1);
      }
      if (tid < 13) sm.Xv[tid] = 0.0;
      __syncthreads();
      RP(4);  // w, g, b phases
      {
        // X_k+1 = L_k X_k + b_k (X_0 = 0), same three sweeps forwards
        constexpr int g = kRicGroup;
        const int k0 = g * sg;  // first step of this slot's group
        if (slot_on) sm.Xv[(k0 + 1) * kVS + sl] = sm.tv[k0 * 13 + sl];
        if (2 * warp < kG)
          ric_chain<1>(&sm.Lk[k0 + 1][sl * 13], 169, &sm.Xv[(k0 + 1) * kVS], kVS, &sm.Xv[(k0 + 2) * kVS + sl], kVS,
                    &sm.tv[(k0 + 1) * 13 + sl], 13, g - 1, slot_on);
        __syncthreads();
        if (warp == 0 && kG > 1) {
          const int l0 = lane < 13 ? lane : 0;
          ric_chain<1>(&Phi[1][l0 * 13], 169, &sm.Xv[g * kVS], g * kVS, &sm.Xv[2 * g * kVS + l0], g * kVS,
                    &sm.Xv[2 * g * kVS + l0], g * kVS, kG - 1, lane < 13);
        }
        __syncthreads();
        if (2 * warp < kG)
          ric_chain<1>(&sm.Lk[k0][sl * 13], 169, &sm.Xv[k0 * kVS], kVS, &sm.Xv[(k0 + 1) * kVS + sl], kVS,
                    &sm.tv[k0 * 13 + sl], 13, g - 1, slot_on);
      }
      RP(5);  // forward chain
      __syncthreads();
      // x~_k = -K_k X_k + g_k ; x <- alpha x~ + (1 - alpha) x
      for (int idx = tid; idx < H * 6; idx += kRicThreads) {  // two outputs per thread
        const int k = idx / 6, a = 2 * (idx - 6 * k);
        double o0, o1;
        ric_pair13(&sm.Kk[k][a], &sm.Xv[k * kVS], o0, o1);
        const double2 g2 = *reinterpret_cast<const double2*>(&sm.gv[12 * k + a]);
        const double2 xo = *reinterpret_cast<const double2*>(&sm.x[12 * k + a]);
        const double xa = g2.x - o0, xb = g2.y - o1;
        *reinterpret_cast<double2*>(&sm.xt[12 * k + a]) = make_double2(xa, xb);
        *reinterpret_cast<double2*>(&sm.x[12 * k + a]) =
            make_double2(alpha * xa + (1.0 - alpha) * xo.x, alpha * xb + (1.0 - alpha) * xo.y);
      }
      __syncthreads();
      // z~ = A x~, z / y updates and the next rhs = sigma x - q + A'(rho z - y), fused: one thread per
      // leg-step owns its five constraint rows and three variables, so no barrier separates the two
      for (int ls = tid; ls < n / 3; ls += kRicThreads) {
        const int r0 = 5 * ls, j0 = 3 * ls;
        const double xtx = sm.xt[j0], xty = sm.xt[j0 + 1], xtz = sm.xt[j0 + 2];
        double wgt[5], ca[5], cz[5];
#pragma unroll
        for (int rw = 0; rw < 5; ++rw) {
          const int i = r0 + rw;
          ca[rw] = sm.cca[i];
          cz[rw] = sm.ccz[i];
          const double zt = ca[rw] * ((rw < 2) ? xtx : xty) + cz[rw] * xtz;
          const double zr = alpha * zt + (1.0 - alpha) * sm.z[i];
          // 1 / rho_i without a division: rho_i is one of three values (auxil.c set_rho_vec)
          const double rvi = sm.rv[i];
          const double rinv = (rvi == rho_cur) ? rinv_in : (rvi == 1e-6) ? rinv_free : rinv_eq;
          const double yi = sm.y[i];
          double zn = zr + rinv * yi;
          const double E = sm.Ev[i];
          const double lo = (double)sm.lb[i] * E, hi = (double)sm.ub[i] * E;
          zn = (zn < lo) ? lo : zn;
          zn = (zn > hi) ? hi : zn;
          const double yn = yi + rvi * (zr - zn);
          sm.y[i] = yn;
          sm.z[i] = zn;
          wgt[rw] = rvi * zn - yn;
        }
        const double sx = ca[0] * wgt[0] + ca[1] * wgt[1];
        const double sy = ca[2] * wgt[2] + ca[3] * wgt[3];
        double sz = 0.0;
#pragma unroll
        for (int rw = 0; rw < 5; ++rw) sz = fma(cz[rw], wgt[rw], sz);
        sm.rhs[j0] = sigma * sm.x[j0] - sm.qb[j0] + sx;
        sm.rhs[j0 + 1] = sigma * sm.x[j0 + 1] - sm.qb[j0 + 1] + sy;
        sm.rhs[j0 + 2] = sigma * sm.x[j0 + 2] - sm.qb[j0 + 2] + sz;
      }
      // next rhs = sigma x - q + A'(rho z - y)
      auto build_rhs = [&]() {
        for (int j = tid; j < n; j += kRicThreads) {
          const int ls = j / 3, vc = j - 3 * ls;
          const int r0 = 5 * ls;
          double s = 0.0;
          if (vc == 0) {
            s = sm.cca[r0] * (sm.rv[r0] * sm.z[r0] - sm.y[r0]) + sm.cca[r0 + 1] * (sm.rv[r0 + 1] * sm.z[r0 + 1] - sm.y[r0 + 1]);
          } else if (vc == 1) {
            s = sm.cca[r0 + 2] * (sm.rv[r0 + 2] * sm.z[r0 + 2] - sm.y[r0 + 2]) +
                sm.cca[r0 + 3] * (sm.rv[r0 + 3] * sm.z[r0 + 3] - sm.y[r0 + 3]);
          } else {
#pragma unroll
            for (int rw = 0; rw < 5; ++rw) s = fma(sm.ccz[r0 + rw], sm.rv[r0 + rw] * sm.z[r0 + rw] - sm.y[r0 + rw], s);
          }
          sm.rhs[j] = sigma * sm.x[j] - sm.qb[j] + s;
        }
      };
      const bool can_check = (--until_check == 0);
      const bool can_adapt = (--until_adapt == 0);
      if (can_check) until_check = sp.check_termination;
      if (can_adapt) until_adapt = sp.adaptive_rho_interval;
      const bool last = (iter == sp.max_iter);
      __syncthreads();
      RP(6);  // x~, z/y, rhs phases
      if (!(can_check || can_adapt || last)) continue;

      // ---- residuals: P x through the open-loop recursions ----
      for (int idx = tid; idx < H * 13; idx += kRicThreads) {  // Bs x_k
        const int k = idx / 13, i = idx - 13 * k;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 12; ++a) s = fma(sm.Bs[k][i * 12 + a], sm.x[12 * k + a], s);
        sm.tv[idx] = s;
      }
      if (tid < kVS) { sm.Xv[tid] = 0.0; sm.pv[(H + 1) * kVS + tid] = 0.0; }
      __syncthreads();
      {
        // X_k+1 = A X_k + (Bs x)_k from X_0 = 0, then mu_i = A' mu_i+1 + c Q X_i for i = H .. 1 (mu_H+1 = 0,
        // stored in pv[i]): the same three sweeps as the closed-loop recursions, with A and A^5
        constexpr int g = kRicGroup;
        const int k0 = g * sg, l0 = lane < 13 ? lane : 0;
        if (slot_on) sm.Xv[(k0 + 1) * kVS + sl] = sm.tv[k0 * 13 + sl];
        if (2 * warp < kG)
          ric_chain<1>(&sm.A[sl * 13], 0, &sm.Xv[(k0 + 1) * kVS], kVS, &sm.Xv[(k0 + 2) * kVS + sl], kVS,
                       &sm.tv[(k0 + 1) * 13 + sl], 13, g - 1, slot_on);
        __syncthreads();
        if (warp == 0 && kG > 1)
          ric_chain<1>(&sm.A5[l0 * 13], 0, &sm.Xv[g * kVS], g * kVS, &sm.Xv[2 * g * kVS + l0], g * kVS,
                       &sm.Xv[2 * g * kVS + l0], g * kVS, kG - 1, lane < 13);
        __syncthreads();
        if (2 * warp < kG)
          ric_chain<1>(&sm.A[sl * 13], 0, &sm.Xv[k0 * kVS], kVS, &sm.Xv[(k0 + 1) * kVS + sl], kVS,
                       &sm.tv[k0 * 13 + sl], 13, g - 1, slot_on);
        __syncthreads();
        for (int idx = tid; idx < H * 13; idx += kRicThreads) {  // addends c Q X_i, i = 1 .. H, at tv[13 (i - 1)]
          const int k = idx / 13, i = idx - 13 * k;
          sm.tv[idx] = c * bp.Qd[i] * sm.Xv[(k + 1) * kVS + i];
        }
        __syncthreads();
        const int il = k0 + g;  // last index of this slot's group
        if (slot_on) sm.pv[il * kVS + sl] = sm.tv[(il - 1) * 13 + sl];
        if (2 * warp < kG)
          ric_chain<13>(&sm.A[sl], 0, &sm.pv[il * kVS], -kVS, &sm.pv[(il - 1) * kVS + sl], -kVS,
                        &sm.tv[(il - 2) * 13 + sl], -13, g - 1, slot_on);
        __syncthreads();
        if (warp == 0 && kG > 1)
          ric_chain<13>(&sm.A5[l0], 0, &sm.pv[(g * (kG - 1) + 1) * kVS], -g * kVS,
                        &sm.pv[(g * (kG - 2) + 1) * kVS + l0], -g * kVS, &sm.pv[(g * (kG - 2) + 1) * kVS + l0],
                        -g * kVS, kG - 1, lane < 13);
        __syncthreads();
        if (2 * warp < kG)
          ric_chain<13>(&sm.A[sl], 0, &sm.pv[(il + 1) * kVS], -kVS, &sm.pv[il * kVS + sl], -kVS,
                        &sm.tv[(il - 1) * 13 + sl], -13, g - 1, slot_on);
      }
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      for (int idx = tid; idx < n; idx += kRicThreads) {
        const int k = idx / 12, a = idx - 12 * k, ls = idx / 3, vc = idx - 3 * ls, r0 = 5 * ls;
        const double d = sm.Dv[idx];
        double px = c * d * bp.Rd[a] * d * sm.x[idx];
#pragma unroll
        for (int i = 0; i < 13; ++i) px = fma(sm.Bs[k][i * 12 + a], sm.pv[(k + 1) * kVS + i], px);
        double aty;
        if (vc == 0) aty = sm.cca[r0] * sm.y[r0] + sm.cca[r0 + 1] * sm.y[r0 + 1];
        else if (vc == 1) aty = sm.cca[r0 + 2] * sm.y[r0 + 2] + sm.cca[r0 + 3] * sm.y[r0 + 3];
        else {
          aty = 0.0;
#pragma unroll
          for (int rw = 0; rw < 5; ++rw) aty = fma(sm.ccz[r0 + rw], sm.y[r0 + rw], aty);
        }
        const double qb = sm.qb[idx], dinv = 1.0 / d;
        const double rd = px + qb + aty;
        v[6] = fmax(v[6], fabs(rd));
        v[7] = fmax(v[7], fabs(dinv * rd));
        v[8] = fmax(v[8], fmax(fmax(fabs(dinv * qb), fabs(dinv * aty)), fabs(dinv * px)));
        v[9] = fmax(v[9], fmax(fmax(fabs(qb), fabs(aty)), fabs(px)));
      }
      for (int i = tid; i < m; i += kRicThreads) {
        const int ls = i / 5, pos = i - 5 * ls;
        const double xl = sm.x[3 * ls + ((pos < 2) ? 0 : 1)], xz = sm.x[3 * ls + 2];
        const double Ax = sm.cca[i] * xl + sm.ccz[i] * xz, zz = sm.z[i], einv = 1.0 / sm.Ev[i];
        const double rp = Ax - zz;
        v[0] = fmax(v[0], fabs(rp));
        v[1] = fmax(v[1], fabs(einv * rp));
        v[2] = fmax(v[2], fabs(einv * zz));
        v[3] = fmax(v[3], fabs(einv * Ax));
        v[4] = fmax(v[4], fabs(zz));
        v[5] = fmax(v[5], fabs(Ax));
      }
      double mres[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) mres[i] = ric_block_reduce<true>(v[i], sm.red, i);
      if (tid == 0) {
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
      RP(7);  // residual check
      if (sm.flags[0]) {
        status = sm.flags[1];
        pri_res_out = sm.scal[4];
        break;
      }
      if (sm.flags[2]) {
        ++rho_updates;
        const double rho = sm.scal[2];
        rho_cur = rho;
        rinv_in = 1.0 / rho;
        rinv_eq = 1.0 / (1e3 * rho);
        for (int i = tid; i < m; i += kRicThreads) {
          const double lo = (double)sm.lb[i] * sm.Ev[i], hi = (double)sm.ub[i] * sm.Ev[i];
          int ctype = 0;
          if (lo < -MPC_INFTY * 1e-4 && hi > MPC_INFTY * 1e-4) ctype = -1;
          else if (hi - lo < 1e-4) ctype = 1;
          const double rvn = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho : rho;
          sm.rv[i] = rvn;
        }
        __syncthreads();
        build_rhs();  // the rhs was built with the old rho vector
        __syncthreads();
        need_factor = true;
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    if (warm) {
      // keep the solver alive for the next tick; a solve that ended with non-finite iterates leaves its
      // slot dead (the robot's next tick is an initSolver), so one bad record cannot poison later ticks
      bool own_ok = true;
      for (int j = tid; j < n; j += kRicThreads) {
        const double xv = sm.x[j];
        own_ok = own_ok && isfinite(xv);
        ws[kWX + j] = xv;
        ws[kWQ + j] = q_all[size_t(p) * n + j];
      }
      for (int i = tid; i < m; i += kRicThreads) {
        const double zv = sm.z[i], yv = sm.y[i];
        own_ok = own_ok && isfinite(zv) && isfinite(yv);
        ws[kWZ + i] = zv;
        ws[kWY + i] = yv;
      }
      const int all_ok = __syncthreads_and(own_ok);
      if (tid == 0) {
        ws[kWRho] = sm.scal[2];
        ws[kWLive] = (all_ok && isfinite(sm.scal[2])) ? 1.0 : 0.0;
      }
    }
    // ---- unscale, rotate the first step to the body frame, write ----
    if (x_all != nullptr)
      for (int j = tid; j < n; j += kRicThreads) x_all[size_t(p) * n + j] = (float)(sm.Dv[j] * sm.x[j]);
    if (tid < 12) {
      const int leg = tid / 3, vc = tid - 3 * leg;
      const double f0 = sm.Dv[3 * leg] * sm.x[3 * leg], f1 = sm.Dv[3 * leg + 1] * sm.x[3 * leg + 1],
                   f2 = sm.Dv[3 * leg + 2] * sm.x[3 * leg + 2];
      double g;
      if (states != nullptr) {
        const float* R = reinterpret_cast<const float*>(states + p) + kOffRot;
        g = (double)R[vc] * f0 + (double)R[3 + vc] * f1 + (double)R[6 + vc] * f2;
      } else {
        g = (vc == 0) ? f0 : (vc == 1) ? f1 : f2;
      }
      const bool bad = isnan(f0) || isnan(f1) || isnan(f2);
      results[p].grf[tid] = bad ? 0.0f : (float)g;
    }
#ifdef RIC_PROF
    if (tid == 0 && p < 2)
      printf("RICPROF p %d iters %d: ruiz %lld factor+ %lld t %lld bwd %lld wgb %lld fwd %lld xzy %lld check %lld\n", p, iter,
             pc_[0], pc_[1], pc_[2], pc_[3], pc_[4], pc_[5], pc_[6], pc_[7]);
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
