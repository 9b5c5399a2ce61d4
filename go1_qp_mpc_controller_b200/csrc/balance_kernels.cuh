// Stance-balance GRF QP (A1RobotControl.cpp:11-48, :321-332, :377-444): 12
// variables, 20 constraints.  ONE WARP PER PROBLEM: lane j < 12 owns variable j
// and row j of P and of -K^-1 (registers), lane i < 20 owns constraint row i.
// All exchanges are warp shuffles; no shared memory, no block barriers.  Same
// OSQP-equivalent ADMM as the MPC path (Ruiz, per-row rho, sweep inverse,
// residual termination, rho adaptation), computed in f64.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpc_b200.h"

namespace mpcb200 {

// Persistent warps: every warp pulls the next problem from an atomic counter.  Iteration counts run from
// 25 to several thousand (mean ~230 at eps 1e-5), so a static warp <-> problem assignment leaves the seven
// other warps of a CTA idle behind one straggler (round 1: grid = n / 8, one 256-thread CTA per SM at 201
// registers = 12 % occupancy).  128 threads x 4 CTAs per SM at <= 128 registers: 16 resident warps per SM.
constexpr int kBalanceThreads = 128;
constexpr int kBalanceCtasPerSm = 4;

struct BalanceParams {
  double Q[6];
  double R, mu, F_min, F_max, mass;
  double kp_lin[3], kd_lin[3], kp_ang[3], kd_ang[3];
  double rho, sigma, alpha, eps_abs, eps_rel, adaptive_rho_tolerance;
  int max_iter, check_termination, scaling, adaptive_rho, adaptive_rho_interval;
};

inline BalanceParams make_balance_params(const BalanceConfig& c) {
  BalanceParams p{};
  for (int i = 0; i < 6; ++i) p.Q[i] = c.Q[i];
  p.R = c.R; p.mu = c.mu; p.F_min = c.F_min; p.F_max = c.F_max; p.mass = c.mass;
  for (int i = 0; i < 3; ++i) {
    p.kp_lin[i] = c.kp_linear[i]; p.kd_lin[i] = c.kd_linear[i];
    p.kp_ang[i] = c.kp_angular[i]; p.kd_ang[i] = c.kd_angular[i];
  }
  p.rho = c.osqp.rho; p.sigma = c.osqp.sigma; p.alpha = c.osqp.alpha;
  p.eps_abs = c.osqp.eps_abs; p.eps_rel = c.osqp.eps_rel;
  p.adaptive_rho_tolerance = c.osqp.adaptive_rho_tolerance;
  p.max_iter = c.osqp.max_iter; p.check_termination = c.osqp.check_termination;
  p.scaling = c.osqp.scaling; p.adaptive_rho = c.osqp.adaptive_rho;
  p.adaptive_rho_interval = c.osqp.adaptive_rho_interval;
  return p;
}

__device__ __forceinline__ double bshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double bwarp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double bwarp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double blimit(double v) {
  v = v < 1e-4 ? 1.0 : v;
  return v > 1e4 ? 1e4 : v;
}

// offsets (floats) inside BalanceStateIn
constexpr int kBEuler = 0, kBPos = 3, kBAngVel = 6, kBLinVel = 9, kBEulerD = 12, kBPosD = 15,
              kBLinVelD = 18, kBAngVelD = 21, kBRot = 24, kBRotZ = 33, kBFoot = 42, kBContacts = 54;

__global__ void __launch_bounds__(kBalanceThreads, kBalanceCtasPerSm)
balance_qp_kernel(const BalanceStateIn* __restrict__ states, int num, int* __restrict__ counter,
                  float* __restrict__ P_out, float* __restrict__ q_out, float* __restrict__ l_out,
                  float* __restrict__ u_out, MpcResult* __restrict__ results,
                  const __grid_constant__ BalanceParams bp) {
  const int lane = threadIdx.x & 31;
 for (;;) {
  int p = 0;
  if (lane == 0) p = atomicAdd(counter, 1);
  p = __shfl_sync(0xffffffffu, p, 0);
  if (p >= num) break;  // warp-uniform
  const float* st = reinterpret_cast<const float*>(states + p);

  // ---- lane roles ----
  const int j = lane < 12 ? lane : 11;   // variable (clamped for idle lanes)
  const int lv = j / 3, cj = j % 3;
  const int i = lane < 20 ? lane : 19;   // constraint row (clamped)
  const bool is_var = lane < 12, is_row = lane < 20;
  const int li = (i < 4) ? i : (i - 4) / 4;        // leg of row i
  const int ki = (i < 4) ? 0 : (i - 4) % 4;
  const int lat = (i < 4) ? -1 : (ki < 2 ? 0 : 1);  // lateral component of the row
  const int ja = 3 * li + (lat < 0 ? 0 : lat);
  const int jz = 3 * li + 2;
  const double ca0 = (i < 4) ? 0.0 : ((ki & 1) ? -1.0 : 1.0);  // rows :33-47
  const double cz0 = (i < 4) ? 1.0 : -bp.mu;
  // rows touching variable j, with unscaled coefficients
  int ridx[5];
  double rco[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) { ridx[t] = 0; rco[t] = 0.0; }
  if (cj == 0) { ridx[0] = 4 + 4 * lv; rco[0] = 1.0; ridx[1] = 5 + 4 * lv; rco[1] = -1.0; }
  else if (cj == 1) { ridx[0] = 6 + 4 * lv; rco[0] = 1.0; ridx[1] = 7 + 4 * lv; rco[1] = -1.0; }
  else {
    ridx[0] = lv; rco[0] = 1.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) { ridx[1 + t] = 4 + 4 * lv + t; rco[1 + t] = -bp.mu; }
  }

  // ---- build: root_acc, M = inertia_inv (6x12), P, q (A1RobotControl.cpp:379-406) ----
  double R[9], Rz[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = (double)st[kBRot + k]; Rz[k] = (double)st[kBRotZ + k]; }
  double ee[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) ee[k] = (double)st[kBEulerD + k] - (double)st[kBEuler + k];
  if (ee[2] > 3.1415926 * 1.5) ee[2] = (double)st[kBEulerD + 2] - 3.1415926 * 2 - (double)st[kBEuler + 2];
  else if (ee[2] < -3.1415926 * 1.5) ee[2] = (double)st[kBEulerD + 2] + 3.1415926 * 2 - (double)st[kBEuler + 2];
  double acc[6];
  {
    double Rtv[3], Rtw[3], tmp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Rtv[k] = R[k] * (double)st[kBLinVel] + R[3 + k] * (double)st[kBLinVel + 1] + R[6 + k] * (double)st[kBLinVel + 2];
      Rtw[k] = R[k] * (double)st[kBAngVel] + R[3 + k] * (double)st[kBAngVel + 1] + R[6 + k] * (double)st[kBAngVel + 2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) tmp[k] = bp.kd_lin[k] * ((double)st[kBLinVelD + k] - Rtv[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      acc[k] = bp.kp_lin[k] * ((double)st[kBPosD + k] - (double)st[kBPos + k]) +
               (R[3 * k] * tmp[0] + R[3 * k + 1] * tmp[1] + R[3 * k + 2] * tmp[2]);
      acc[3 + k] = bp.kp_ang[k] * ee[k] + bp.kd_ang[k] * ((double)st[kBAngVelD + k] - Rtw[k]);
    }
    acc[2] += bp.mass * 9.8;
  }
  // column j of M: top = e_cj, bottom = column cj of Rz' * skew(foot_lv)
  double mj[6];
  {
    const double fx = st[kBFoot + 3 * lv], fy = st[kBFoot + 3 * lv + 1], fz = st[kBFoot + 3 * lv + 2];
    const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      mj[k] = (k == cj) ? 1.0 : 0.0;
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < 3; ++t) s += Rz[3 * t + k] * sk[3 * t + cj];  // (Rz')[k][t] = Rz[t][k]
      mj[3 + k] = s;
    }
  }
  double Prow[12];  // row j of P
  {
    double qm[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) qm[k] = mj[k] * bp.Q[k];
#pragma unroll
    for (int b = 0; b < 12; ++b) {
      double s = (b == j) ? bp.R : 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += qm[k] * bshfl(mj[k], b);
      Prow[b] = s;
    }
  }
  double q0;
  {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += mj[k] * bp.Q[k] * acc[k];
    q0 = -s;
  }
  // bounds (:409-413, :33-47)
  double lb, ub;
  if (i < 4) {
    const double cf = (st[kBContacts + i] != 0.0f) ? 1.0 : 0.0;
    lb = cf * bp.F_min;
    ub = cf * bp.F_max;
  } else {
    lb = -MPC_INFTY;
    ub = 0.0;
  }
  if (P_out != nullptr && is_var) {
#pragma unroll
    for (int b = 0; b < 12; ++b) P_out[size_t(p) * 144 + j * 12 + b] = (float)Prow[b];
  }
  if (q_out != nullptr && is_var) q_out[size_t(p) * 12 + j] = (float)q0;
  if (l_out != nullptr && is_row) {
    l_out[size_t(p) * 20 + i] = (float)lb;
    u_out[size_t(p) * 20 + i] = (float)ub;
  }
  // (the QP never leaves registers here, so unlike the MPC path it is NOT
  // rounded to fp32 before the solve; the fp32 copies above are parity output)

  // ---- Ruiz equilibration (lane j: D_j, lane i: E_i) ----
  double D = 1.0, E = 1.0, c = 1.0;
  auto p_row_norm = [&](double Dj) {
    double m = 0.0;
#pragma unroll
    for (int b = 0; b < 12; ++b) m = fmax(m, fabs(Prow[b]) * bshfl(Dj, b));
    return m;
  };
  if (bp.scaling > 0) {
    double nP = p_row_norm(D);
    for (int it = 0; it < bp.scaling; ++it) {
      // column norm of A for variable j
      double nA = 0.0;
#pragma unroll
      for (int t = 0; t < 5; ++t) nA = fmax(nA, fabs(rco[t]) * bshfl(E, ridx[t]));
      nA *= D;
      const double Dt = is_var ? rsqrt(blimit(fmax(nP, nA))) : 1.0;
      // row norm of A for constraint i
      const double da = bshfl(D, ja), dz = bshfl(D, jz);
      const double nrow = E * fmax(fabs(ca0) * da, fabs(cz0) * dz);
      const double Et = is_row ? rsqrt(blimit(nrow)) : 1.0;
      D *= Dt;
      E *= Et;
      const double nP2 = c * D * p_row_norm(D);
      const double mean = bwarp_sum(is_var ? nP2 : 0.0) / 12.0;
      const double qn = bwarp_max(is_var ? fabs(c * D * q0) : 0.0);
      const double ct = 1.0 / blimit(fmax(mean, blimit(qn)));
      c *= ct;
      nP = nP2 * ct;
    }
  }
  const double cinv = 1.0 / c, Dinv = 1.0 / D, Einv = 1.0 / E;
  const double qb = c * D * q0;
  lb *= E;
  ub *= E;
  const int ctype = (lb < -MPC_INFTY * 1e-4 && ub > MPC_INFTY * 1e-4) ? -1 : ((ub - lb < 1e-4) ? 1 : 0);
  double rho = bp.rho;
  auto rho_of = [&](double rh) { return (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rh : rh; };
  double rv = rho_of(rho), rinv = 1.0 / rv;
  // scaled constraint coefficients of row i
  const double ca = E * ca0 * bshfl(D, ja);
  const double cz = E * cz0 * bshfl(D, jz);
  // scaled coefficients of the rows touching variable j
  double rcs[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) rcs[t] = bshfl(E, ridx[t]) * rco[t] * D;

  double a[12];  // row j of -K^-1
  auto factor = [&]() {
    // K = c D P D + sigma I + A' diag(rho) A
#pragma unroll
    for (int b = 0; b < 12; ++b) {
      double v = c * D * Prow[b] * bshfl(D, b);
      if (b == j) v += bp.sigma;
      a[b] = v;
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int ri = ridx[t];
      const double rr = bshfl(rv, ri);
      const double ca_i = bshfl(ca, ri), cz_i = bshfl(cz, ri);
      const int lat_i = (ri < 4) ? -1 : (((ri - 4) % 4) < 2 ? 0 : 1);
      const double wgt = rr * rcs[t];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const double aib = (cc == 2) ? cz_i : ((cc == lat_i) ? ca_i : 0.0);
        const int b = 3 * lv + cc;
#pragma unroll
        for (int bb = 0; bb < 12; ++bb)
          if (bb == b) a[bb] += wgt * aib;
      }
    }
    // symmetric sweep, 12 pivots
#pragma unroll 1
    for (int k = 0; k < 12; ++k) {
      double v[12];
#pragma unroll
      for (int b = 0; b < 12; ++b) v[b] = bshfl(a[b], k);
      double d = 0.0;
#pragma unroll
      for (int b = 0; b < 12; ++b)
        if (b == k) d = v[b];
      const double dinv = 1.0 / d;
      double vr = 0.0;
#pragma unroll
      for (int b = 0; b < 12; ++b)
        if (b == j) vr = v[b];
      if (j != k) {
        const double w = -vr * dinv;
#pragma unroll
        for (int b = 0; b < 12; ++b) {
          const double vb = (b == k) ? (d - 1.0) : v[b];
          a[b] = fma(w, vb, a[b]);
        }
      } else {
#pragma unroll
        for (int b = 0; b < 12; ++b) a[b] = (b == k) ? -dinv : a[b] * dinv;
      }
    }
  };
  factor();

  // ---- ADMM ----
  double x = 0.0, z = 0.0, y = 0.0;
  int iter = 0, status = MPC_STATUS_UNSOLVED, rho_updates = 0;
  double pri_out = 0.0;
  // iterations run in stretches up to the next event (termination check, rho adaptation, iteration
  // limit) on a plain counter: the two runtime `%` of the straightforward loop were a third of its
  // instructions (ncu r02_balance_v2: 320 warp instructions per iteration of a 12-variable problem)
  const int chk = bp.check_termination > 0 ? bp.check_termination : 0x7fffffff;
  const int adp = (bp.adaptive_rho && bp.adaptive_rho_interval > 0) ? bp.adaptive_rho_interval : 0x7fffffff;
  int until_check = chk, until_adapt = adp;
  iter = 0;
  for (;;) {
    int run = until_check < until_adapt ? until_check : until_adapt;
    run = run < bp.max_iter - iter ? run : bp.max_iter - iter;
#pragma unroll 1
    for (int qq = 0; qq < run; ++qq) {
      // rhs_j = sigma x - q + A'(rho z - y)
      const double w = rv * z - y;
      double rhs = bp.sigma * x - qb;
#pragma unroll
      for (int t = 0; t < 5; ++t) rhs = fma(rcs[t], bshfl(w, ridx[t]), rhs);
      double xt0 = 0.0, xt1 = 0.0;
#pragma unroll
      for (int b = 0; b < 12; b += 2) {
        xt0 = fma(a[b], bshfl(rhs, b), xt0);
        xt1 = fma(a[b + 1], bshfl(rhs, b + 1), xt1);
      }
      const double xt = -(xt0 + xt1);
      x = bp.alpha * xt + (1.0 - bp.alpha) * x;
      const double zt = ca * bshfl(xt, ja) + cz * bshfl(xt, jz);
      const double zr = bp.alpha * zt + (1.0 - bp.alpha) * z;
      double zn = zr + rinv * y;
      zn = (zn < lb) ? lb : zn;
      zn = (zn > ub) ? ub : zn;
      y += rv * (zr - zn);
      z = zn;
    }
    iter += run;
    until_check -= run;
    until_adapt -= run;
    const bool can_check = until_check == 0, can_adapt = until_adapt == 0;
    if (can_check) until_check = chk;
    if (can_adapt) until_adapt = adp;
    const bool last = iter == bp.max_iter;
    // residuals
    const double Ax = ca * bshfl(x, ja) + cz * bshfl(x, jz);
    const double rp = is_row ? (Ax - z) : 0.0;
    double Px = 0.0;
    {
      const double xD = D * x;
#pragma unroll
      for (int b = 0; b < 12; ++b) Px = fma(Prow[b], bshfl(xD, b), Px);
      Px *= c * D;
    }
    double Aty = 0.0;
#pragma unroll
    for (int t = 0; t < 5; ++t) Aty = fma(rcs[t], bshfl(y, ridx[t]), Aty);
    const double rd = is_var ? (Px + qb + Aty) : 0.0;
    const double m0 = bwarp_max(fabs(rp));
    const double m1 = bwarp_max(fabs(Einv * rp));
    const double m2 = bwarp_max(is_row ? fmax(fabs(Einv * z), fabs(Einv * Ax)) : 0.0);
    const double m4 = bwarp_max(is_row ? fmax(fabs(z), fabs(Ax)) : 0.0);
    const double m6 = bwarp_max(fabs(rd));
    const double m7 = bwarp_max(fabs(Dinv * rd));
    const double m8 = bwarp_max(is_var ? fmax(fmax(fabs(Dinv * qb), fabs(Dinv * Aty)), fabs(Dinv * Px)) : 0.0);
    const double m9 = bwarp_max(is_var ? fmax(fmax(fabs(qb), fabs(Aty)), fabs(Px)) : 0.0);
    const double pri = m1, dua = cinv * m7;
    const double eps_pri = bp.eps_abs + bp.eps_rel * m2;
    const double eps_dua = bp.eps_abs + bp.eps_rel * cinv * m8;
    pri_out = pri;
    if ((can_check || last) && pri < eps_pri && dua < eps_dua) { status = MPC_STATUS_SOLVED; break; }
    if (last) {
      status = (pri < 10.0 * eps_pri && dua < 10.0 * eps_dua) ? 2 : MPC_STATUS_MAX_ITER_REACHED;
      break;
    }
    if (can_adapt) {
      const double pn = m0 / (m4 + 1e-10);
      const double dn = m6 / (m9 + 1e-10);
      double rho_new = rho * sqrt(pn / (dn + 1e-10));
      rho_new = fmin(fmax(rho_new, 1e-6), 1e6);
      if (rho_new > rho * bp.adaptive_rho_tolerance || rho_new < rho / bp.adaptive_rho_tolerance) {
        rho = rho_new;
        rv = rho_of(rho);
        rinv = 1.0 / rv;
        ++rho_updates;
        factor();
      }
    }
  }

  // ---- unscale and rotate every leg to the body frame (:439-444) ----
  const double xo = D * x;
  const int leg = lane / 3 < 4 ? lane / 3 : 3;
  const double f0 = bshfl(xo, 3 * leg), f1 = bshfl(xo, 3 * leg + 1), f2 = bshfl(xo, 3 * leg + 2);
  if (is_var) {
    const double g = R[cj] * f0 + R[3 + cj] * f1 + R[6 + cj] * f2;
    const bool bad = isnan(f0) || isnan(f1) || isnan(f2);
    results[p].grf[j] = bad ? 0.0f : (float)g;
  }
  if (lane == 12) {
    results[p].status = status;
    results[p].iters = iter;
    results[p].rho_updates = rho_updates;
    results[p].pri_res = (float)pri_out;
  }
 }  // next problem
}


// ---------------------------------------------------------------------------------------------------------
// Second layout of the same QP: FOUR LANES PER PROBLEM, eight problems per warp.  The constraint rows of the
// balance QP are leg-local (row i: fz_i; rows 4+4i .. 7+4i: +-fx_i - mu fz_i, +-fy_i - mu fz_i), so a lane that
// owns a leg -- its three variables, its five rows, three rows of P and of -K^-1 -- does everything but the
// 12-term dot products and the norms inside itself; those take the other legs' values by width-4 shuffles.
// The one-warp-per-problem kernel above issues 243 warp instructions per ADMM iteration for ONE problem (12
// to 20 of 32 lanes busy, a 64-bit shuffle pair per exchanged value); this one ~135 for EIGHT.  The eight
// problems of a warp run in lockstep on one iteration counter (checks and rho updates fall on the same
// iterations for everybody, so a refactorisation is one pass with the groups that need it active); a group that
// has converged waits for the slowest of its warp -- with iteration counts from 25 to a few thousand around a
// mean of 230 that costs about half of the lane-time, and still leaves a factor ~5 over the warp kernel.
// Arithmetic is the warp kernel's, operation for operation (same products, same summation order inside the dot
// products, same sweep); only the cross-variable sums of the cost normalisation are ordered differently.
constexpr int kBalLegThreads = 128;
constexpr int kBalLegCtasPerSm = 2;
constexpr int kBalLegPStride = 146;  // doubles per problem in shared memory (144 + 2: groups on different banks)

__device__ __forceinline__ double bshfl4(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 4); }
__device__ __forceinline__ double bmax4(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1, 4));
  return fmax(v, __shfl_xor_sync(0xffffffffu, v, 2, 4));
}
// max of NON-NEGATIVE doubles through their bit patterns: a NaN compares above everything, so a bad state record
// reaches the termination test (and fails it) instead of being dropped by fmax and reported as solved
__device__ __forceinline__ double bmaxn(double a, double b) {
  return (__double_as_longlong(a) > __double_as_longlong(b)) ? a : b;
}
__device__ __forceinline__ double bmaxn4(double v) {
  v = bmaxn(v, __shfl_xor_sync(0xffffffffu, v, 1, 4));
  return bmaxn(v, __shfl_xor_sync(0xffffffffu, v, 2, 4));
}

__global__ void __launch_bounds__(kBalLegThreads, kBalLegCtasPerSm)
balance_qp_leg_kernel(const BalanceStateIn* __restrict__ states, int num, int* __restrict__ counter,
                      float* __restrict__ P_out, float* __restrict__ q_out, float* __restrict__ l_out,
                      float* __restrict__ u_out, MpcResult* __restrict__ results,
                      const __grid_constant__ BalanceParams bp) {
  extern __shared__ __align__(16) double bal_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 2, lg = lane & 3;                 // problem of the warp, leg
  double* const Ps = bal_smem + (warp * 8 + grp) * kBalLegPStride;   // P of this group's problem, 12 x 12
  const double mu = bp.mu;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, 8);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= num) break;  // warp-uniform
    const int p = base + grp;
    const bool valid = p < num;
    const float* st = reinterpret_cast<const float*>(states + (valid ? p : num - 1));

    // ---- build: root_acc, M = inertia_inv (6x12), P, q (A1RobotControl.cpp:379-406) ----
    double R[9], Rz[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { R[k] = (double)st[kBRot + k]; Rz[k] = (double)st[kBRotZ + k]; }
    double acc[6];
    {
      double ee[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) ee[k] = (double)st[kBEulerD + k] - (double)st[kBEuler + k];
      if (ee[2] > 3.1415926 * 1.5) ee[2] = (double)st[kBEulerD + 2] - 3.1415926 * 2 - (double)st[kBEuler + 2];
      else if (ee[2] < -3.1415926 * 1.5) ee[2] = (double)st[kBEulerD + 2] + 3.1415926 * 2 - (double)st[kBEuler + 2];
      double Rtv[3], Rtw[3], tmp[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Rtv[k] = R[k] * (double)st[kBLinVel] + R[3 + k] * (double)st[kBLinVel + 1] + R[6 + k] * (double)st[kBLinVel + 2];
        Rtw[k] = R[k] * (double)st[kBAngVel] + R[3 + k] * (double)st[kBAngVel + 1] + R[6 + k] * (double)st[kBAngVel + 2];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) tmp[k] = bp.kd_lin[k] * ((double)st[kBLinVelD + k] - Rtv[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        acc[k] = bp.kp_lin[k] * ((double)st[kBPosD + k] - (double)st[kBPos + k]) +
                 (R[3 * k] * tmp[0] + R[3 * k + 1] * tmp[1] + R[3 * k + 2] * tmp[2]);
        acc[3 + k] = bp.kp_ang[k] * ee[k] + bp.kd_ang[k] * ((double)st[kBAngVelD + k] - Rtw[k]);
      }
      acc[2] += bp.mass * 9.8;
    }
    // bottom halves of the leg's three columns of M: column cj of Rz' * skew(foot) (the top halves are e_cj)
    double mb[3][3];  // mb[cj][k] = M[3 + k][3 lg + cj]
    {
      const double fx = st[kBFoot + 3 * lg], fy = st[kBFoot + 3 * lg + 1], fz = st[kBFoot + 3 * lg + 2];
      const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};
#pragma unroll
      for (int cj = 0; cj < 3; ++cj)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < 3; ++t) s += Rz[3 * t + k] * sk[3 * t + cj];  // (Rz')[k][t] = Rz[t][k]
          mb[cj][k] = s;
        }
    }
    // rows 3 lg .. 3 lg + 2 of P:  P[j][b] = sum_k (m_j[k] Q_k) m_b[k] + R [b == j], k in the warp kernel's order
    double Prow[3][12], q0[3];
#pragma unroll
    for (int lb = 0; lb < 4; ++lb)
#pragma unroll
      for (int cb = 0; cb < 3; ++cb) {
        const double o0 = bshfl4(mb[cb][0], lb), o1 = bshfl4(mb[cb][1], lb), o2 = bshfl4(mb[cb][2], lb);
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          double s = (lb == lg && cb == cj) ? bp.R : 0.0;
          // k = 0..2: m_j[k] = [k == cj], m_b[k] = [k == cb]; k = 3..5: bottom halves
          if (cj == cb) s += (1.0 * bp.Q[cj]) * 1.0;
          s += (mb[cj][0] * bp.Q[3]) * o0;
          s += (mb[cj][1] * bp.Q[4]) * o1;
          s += (mb[cj][2] * bp.Q[5]) * o2;
          Prow[cj][3 * lb + cb] = s;
        }
      }
#pragma unroll
    for (int cj = 0; cj < 3; ++cj) {
      double s = 0.0;
      s += 1.0 * bp.Q[cj] * acc[cj];
      s += mb[cj][0] * bp.Q[3] * acc[3];
      s += mb[cj][1] * bp.Q[4] * acc[4];
      s += mb[cj][2] * bp.Q[5] * acc[5];
      q0[cj] = -s;
    }
    // bounds (:409-413, :33-47): row 0 = fz of the leg, rows 1..4 = +fx, -fx, +fy, -fy - mu fz in (-inf, 0]
    const double cf = (st[kBContacts + lg] != 0.0f) ? 1.0 : 0.0;
    double lb0 = cf * bp.F_min, ub0 = cf * bp.F_max;
    if (valid) {
      if (P_out != nullptr) {
#pragma unroll
        for (int cj = 0; cj < 3; ++cj)
#pragma unroll
          for (int b = 0; b < 12; ++b) P_out[size_t(p) * 144 + (3 * lg + cj) * 12 + b] = (float)Prow[cj][b];
      }
      if (q_out != nullptr) {
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) q_out[size_t(p) * 12 + 3 * lg + cj] = (float)q0[cj];
      }
      if (l_out != nullptr) {
        l_out[size_t(p) * 20 + lg] = (float)lb0;
        u_out[size_t(p) * 20 + lg] = (float)ub0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          l_out[size_t(p) * 20 + 4 + 4 * lg + t] = (float)(-MPC_INFTY);
          u_out[size_t(p) * 20 + 4 + 4 * lg + t] = 0.0f;
        }
      }
    }
    // P to shared memory for the factorisations and the residual checks
    __syncwarp();
#pragma unroll
    for (int cj = 0; cj < 3; ++cj)
#pragma unroll
      for (int b = 0; b < 12; ++b) Ps[(3 * lg + cj) * 12 + b] = Prow[cj][b];
    __syncwarp();

    // ---- Ruiz equilibration: D of the leg's variables, E of its rows (row 0, then +fx, -fx, +fy, -fy) ----
    double D[3] = {1.0, 1.0, 1.0}, E[5] = {1.0, 1.0, 1.0, 1.0, 1.0}, c = 1.0;
    auto p_row_norms = [&](double (&nrm)[3]) {
      nrm[0] = nrm[1] = nrm[2] = 0.0;
#pragma unroll
      for (int lb = 0; lb < 4; ++lb)
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
          const double db = bshfl4(D[cb], lb);
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) nrm[cj] = fmax(nrm[cj], fabs(Prow[cj][3 * lb + cb]) * db);
        }
    };
    if (bp.scaling > 0) {
      double nP[3];
      p_row_norms(nP);
      for (int it = 0; it < bp.scaling; ++it) {
        double nA[3];
        nA[0] = fmax(E[1], E[2]) * D[0];
        nA[1] = fmax(E[3], E[4]) * D[1];
        {
          double m = fabs(1.0) * E[0];
#pragma unroll
          for (int t = 1; t < 5; ++t) m = fmax(m, fabs(-mu) * E[t]);
          nA[2] = m * D[2];
        }
        double Dt[3], Et[5];
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) Dt[cj] = rsqrt(blimit(fmax(nP[cj], nA[cj])));
        Et[0] = rsqrt(blimit(E[0] * fmax(0.0 * D[0], 1.0 * D[2])));
        Et[1] = rsqrt(blimit(E[1] * fmax(1.0 * D[0], fabs(-mu) * D[2])));
        Et[2] = rsqrt(blimit(E[2] * fmax(1.0 * D[0], fabs(-mu) * D[2])));
        Et[3] = rsqrt(blimit(E[3] * fmax(1.0 * D[1], fabs(-mu) * D[2])));
        Et[4] = rsqrt(blimit(E[4] * fmax(1.0 * D[1], fabs(-mu) * D[2])));
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) D[cj] *= Dt[cj];
#pragma unroll
        for (int t = 0; t < 5; ++t) E[t] *= Et[t];
        double nr[3], nP2[3];
        p_row_norms(nr);
        double part = 0.0, qn = 0.0;
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          nP2[cj] = c * D[cj] * nr[cj];
          part += nP2[cj];
          qn = fmax(qn, fabs(c * D[cj] * q0[cj]));
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1, 4);
        part += __shfl_xor_sync(0xffffffffu, part, 2, 4);
        qn = bmax4(qn);
        const double mean = part / 12.0;
        const double ct = 1.0 / blimit(fmax(mean, blimit(qn)));
        c *= ct;
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) nP[cj] = nP2[cj] * ct;
      }
    }
    const double cinv = 1.0 / c;
    double qb[3];
#pragma unroll
    for (int cj = 0; cj < 3; ++cj) qb[cj] = c * D[cj] * q0[cj];
    // scaled bounds, constraint types, rho of the rows
    lb0 *= E[0];
    ub0 *= E[0];
    const int ct0 = (lb0 < -MPC_INFTY * 1e-4 && ub0 > MPC_INFTY * 1e-4) ? -1 : ((ub0 - lb0 < 1e-4) ? 1 : 0);
    int ctp = ct0 + 1;
    double lbr[5], ubr[5];
    lbr[0] = lb0; ubr[0] = ub0;
#pragma unroll
    for (int t = 1; t < 5; ++t) {
      lbr[t] = -MPC_INFTY * E[t];
      ubr[t] = 0.0 * E[t];
      const int ctt = (lbr[t] < -MPC_INFTY * 1e-4 && ubr[t] > MPC_INFTY * 1e-4) ? -1 : ((ubr[t] - lbr[t] < 1e-4) ? 1 : 0);
      ctp |= (ctt + 1) << (2 * t);
    }
    auto rho_row = [&](int t, double rh) {
      const int ctt = ((ctp >> (2 * t)) & 3) - 1;
      return (ctt == -1) ? 1e-6 : (ctt == 1) ? 1e3 * rh : rh;
    };
    double rho = bp.rho;
    double rv[5], rinv[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) { rv[t] = rho_row(t, rho); rinv[t] = 1.0 / rv[t]; }
    // scaled constraint coefficients: row t = ca[t] x_lat(t) + cz[t] x_fz,  lat(1, 2) = fx, lat(3, 4) = fy
    double ca[5], cz[5];
    ca[0] = E[0] * 0.0 * D[0];
    cz[0] = E[0] * 1.0 * D[2];
    ca[1] = E[1] * 1.0 * D[0];  cz[1] = E[1] * (-mu) * D[2];
    ca[2] = E[2] * (-1.0) * D[0]; cz[2] = E[2] * (-mu) * D[2];
    ca[3] = E[3] * 1.0 * D[1];  cz[3] = E[3] * (-mu) * D[2];
    ca[4] = E[4] * (-1.0) * D[1]; cz[4] = E[4] * (-mu) * D[2];

    double a[3][12];  // rows 3 lg .. 3 lg + 2 of -K^-1
    auto factor = [&]() {
      // K = c D P D + sigma I + A' diag(rho) A
#pragma unroll
      for (int lb = 0; lb < 4; ++lb)
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
          const double db = bshfl4(D[cb], lb);
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) {
            double v = c * D[cj] * Ps[(3 * lg + cj) * 12 + 3 * lb + cb] * db;
            if (lb == lg && cb == cj) v += bp.sigma;
            a[cj][3 * lb + cb] = v;
          }
        }
      // A' rho A on the leg's own 3 x 3 block, rows in the warp kernel's order (fx: rows 1, 2; fy: 3, 4; fz: 0, 1, 2, 3, 4)
#pragma unroll
      for (int lb = 0; lb < 4; ++lb) {
        if (lb == lg) {
          // variable fx
          {
            const double w1 = rv[1] * ca[1], w2 = rv[2] * ca[2];
            a[0][3 * lb + 0] += w1 * ca[1]; a[0][3 * lb + 2] += w1 * cz[1];
            a[0][3 * lb + 0] += w2 * ca[2]; a[0][3 * lb + 2] += w2 * cz[2];
          }
          {
            const double w3 = rv[3] * ca[3], w4 = rv[4] * ca[4];
            a[1][3 * lb + 1] += w3 * ca[3]; a[1][3 * lb + 2] += w3 * cz[3];
            a[1][3 * lb + 1] += w4 * ca[4]; a[1][3 * lb + 2] += w4 * cz[4];
          }
          {
            const double w0 = rv[0] * cz[0];
            a[2][3 * lb + 2] += w0 * cz[0];
            const double w1 = rv[1] * cz[1], w2 = rv[2] * cz[2], w3 = rv[3] * cz[3], w4 = rv[4] * cz[4];
            a[2][3 * lb + 0] += w1 * ca[1]; a[2][3 * lb + 2] += w1 * cz[1];
            a[2][3 * lb + 0] += w2 * ca[2]; a[2][3 * lb + 2] += w2 * cz[2];
            a[2][3 * lb + 1] += w3 * ca[3]; a[2][3 * lb + 2] += w3 * cz[3];
            a[2][3 * lb + 1] += w4 * ca[4]; a[2][3 * lb + 2] += w4 * cz[4];
          }
        }
      }
      // symmetric sweep, 12 pivots; pivot row k lives in lane k / 3 of the group
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        double v[12];
#pragma unroll
        for (int b = 0; b < 12; ++b) v[b] = bshfl4(a[k % 3][b], k / 3);
        const double d = v[k];
        const double dinv = 1.0 / d;
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          const bool pivot_row = (k / 3 == lg) && (k % 3 == cj);
          // v[3 lg + cj] without a dynamic index
          double vr = 0.0;
#pragma unroll
          for (int lb = 0; lb < 4; ++lb)
            if (lb == lg) vr = v[3 * lb + cj];
          const double w = -vr * dinv;
#pragma unroll
          for (int b = 0; b < 12; ++b) {
            const double vb = (b == k) ? (d - 1.0) : v[b];
            const double upd = fma(w, vb, a[cj][b]);
            const double prw = (b == k) ? -dinv : a[cj][b] * dinv;
            a[cj][b] = pivot_row ? prw : upd;
          }
        }
      }
    };
    factor();

    // ---- ADMM, the eight problems of the warp in lockstep ----
    double x[3] = {0.0, 0.0, 0.0}, z[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, y[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int iter = 0;
    bool done = !valid;
    int status_out = MPC_STATUS_UNSOLVED, iter_out = 0, rho_updates = 0, rho_out = 0;
    double pri_out = 0.0, xo[3] = {0.0, 0.0, 0.0};
    const int chk = bp.check_termination > 0 ? bp.check_termination : 0x7fffffff;
    const int adp = (bp.adaptive_rho && bp.adaptive_rho_interval > 0) ? bp.adaptive_rho_interval : 0x7fffffff;
    int until_check = chk, until_adapt = adp;
    for (;;) {
      int run = until_check < until_adapt ? until_check : until_adapt;
      run = run < bp.max_iter - iter ? run : bp.max_iter - iter;
#pragma unroll 1
      for (int qq = 0; qq < run; ++qq) {
        // rhs = sigma x - q + A'(rho z - y), rows in the warp kernel's order
        double w[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) w[t] = rv[t] * z[t] - y[t];
        double rhs[3];
        rhs[0] = fma(ca[2], w[2], fma(ca[1], w[1], bp.sigma * x[0] - qb[0]));
        rhs[1] = fma(ca[4], w[4], fma(ca[3], w[3], bp.sigma * x[1] - qb[1]));
        rhs[2] = fma(cz[4], w[4], fma(cz[3], w[3], fma(cz[2], w[2], fma(cz[1], w[1], fma(cz[0], w[0], bp.sigma * x[2] - qb[2])))));
        double e0[3] = {0.0, 0.0, 0.0}, e1[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int lb = 0; lb < 4; ++lb) {
          const double r0 = bshfl4(rhs[0], lb), r1 = bshfl4(rhs[1], lb), r2 = bshfl4(rhs[2], lb);
          // b = 3 lb, 3 lb + 1, 3 lb + 2 go to the even / odd accumulator by the parity of b
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) {
            if ((3 * lb) % 2 == 0) {
              e0[cj] = fma(a[cj][3 * lb], r0, e0[cj]);
              e1[cj] = fma(a[cj][3 * lb + 1], r1, e1[cj]);
              e0[cj] = fma(a[cj][3 * lb + 2], r2, e0[cj]);
            } else {
              e1[cj] = fma(a[cj][3 * lb], r0, e1[cj]);
              e0[cj] = fma(a[cj][3 * lb + 1], r1, e0[cj]);
              e1[cj] = fma(a[cj][3 * lb + 2], r2, e1[cj]);
            }
          }
        }
        double xt[3];
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          xt[cj] = -(e0[cj] + e1[cj]);
          x[cj] = bp.alpha * xt[cj] + (1.0 - bp.alpha) * x[cj];
        }
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const double xl = (t == 0) ? xt[0] : (t < 3 ? xt[0] : xt[1]);
          const double zt = ca[t] * xl + cz[t] * xt[2];
          const double zr = bp.alpha * zt + (1.0 - bp.alpha) * z[t];
          double zn = zr + rinv[t] * y[t];
          zn = (zn < lbr[t]) ? lbr[t] : zn;
          zn = (zn > ubr[t]) ? ubr[t] : zn;
          y[t] += rv[t] * (zr - zn);
          z[t] = zn;
        }
      }
      iter += run;
      until_check -= run;
      until_adapt -= run;
      const bool can_check = until_check == 0, can_adapt = until_adapt == 0;
      if (can_check) until_check = chk;
      if (can_adapt) until_adapt = adp;
      const bool last = iter == bp.max_iter;
      // residuals
      double m0 = 0.0, m1 = 0.0, m2 = 0.0, m4 = 0.0, m6 = 0.0, m7 = 0.0, m8 = 0.0, m9 = 0.0;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const double xl = (t < 3) ? x[0] : x[1];
        const double Ax = ca[t] * xl + cz[t] * x[2];
        const double rp = Ax - z[t];
        const double Einv = 1.0 / E[t];
        m0 = bmaxn(m0, fabs(rp));
        m1 = bmaxn(m1, fabs(Einv * rp));
        m2 = bmaxn(m2, bmaxn(fabs(Einv * z[t]), fabs(Einv * Ax)));
        m4 = bmaxn(m4, bmaxn(fabs(z[t]), fabs(Ax)));
      }
      {
        double Px[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int lb = 0; lb < 4; ++lb)
#pragma unroll
          for (int cb = 0; cb < 3; ++cb) {
            const double xd = bshfl4(D[cb] * x[cb], lb);
#pragma unroll
            for (int cj = 0; cj < 3; ++cj) Px[cj] = fma(Ps[(3 * lg + cj) * 12 + 3 * lb + cb], xd, Px[cj]);
          }
        double Aty[3];
        Aty[0] = fma(ca[2], y[2], fma(ca[1], y[1], 0.0));
        Aty[1] = fma(ca[4], y[4], fma(ca[3], y[3], 0.0));
        Aty[2] = fma(cz[4], y[4], fma(cz[3], y[3], fma(cz[2], y[2], fma(cz[1], y[1], fma(cz[0], y[0], 0.0)))));
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) {
          const double px = Px[cj] * (c * D[cj]);
          const double rd = px + qb[cj] + Aty[cj];
          const double Dinv = 1.0 / D[cj];
          m6 = bmaxn(m6, fabs(rd));
          m7 = bmaxn(m7, fabs(Dinv * rd));
          m8 = bmaxn(m8, bmaxn(bmaxn(fabs(Dinv * qb[cj]), fabs(Dinv * Aty[cj])), fabs(Dinv * px)));
          m9 = bmaxn(m9, bmaxn(bmaxn(fabs(qb[cj]), fabs(Aty[cj])), fabs(px)));
        }
      }
      m0 = bmaxn4(m0); m1 = bmaxn4(m1); m2 = bmaxn4(m2); m4 = bmaxn4(m4);
      m6 = bmaxn4(m6); m7 = bmaxn4(m7); m8 = bmaxn4(m8); m9 = bmaxn4(m9);
      const double pri = m1, dua = cinv * m7;
      const double eps_pri = bp.eps_abs + bp.eps_rel * m2;
      const double eps_dua = bp.eps_abs + bp.eps_rel * cinv * m8;
      bool refactor = false;
      if (!done) {
        bool fin = false;
        int stt = MPC_STATUS_UNSOLVED;
        if ((can_check || last) && pri < eps_pri && dua < eps_dua) { fin = true; stt = MPC_STATUS_SOLVED; }
        else if (last) { fin = true; stt = (pri < 10.0 * eps_pri && dua < 10.0 * eps_dua) ? 2 : MPC_STATUS_MAX_ITER_REACHED; }
        if (fin) {
          done = true;
          status_out = stt; iter_out = iter; rho_out = rho_updates; pri_out = pri;
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) xo[cj] = D[cj] * x[cj];
        } else if (can_adapt) {
          const double pn = m0 / (m4 + 1e-10);
          const double dn = m6 / (m9 + 1e-10);
          double rho_new = rho * sqrt(pn / (dn + 1e-10));
          rho_new = fmin(fmax(rho_new, 1e-6), 1e6);
          if (rho_new > rho * bp.adaptive_rho_tolerance || rho_new < rho / bp.adaptive_rho_tolerance) {
            rho = rho_new;
#pragma unroll
            for (int t = 0; t < 5; ++t) { rv[t] = rho_row(t, rho); rinv[t] = 1.0 / rv[t]; }
            ++rho_updates;
            refactor = true;
          }
        }
      }
      if (__all_sync(0xffffffffu, done)) break;
      // one pass of the factorisation; groups that do not need it recompute the factors they have (bit-identical)
      if (__any_sync(0xffffffffu, refactor)) factor();
    }

    // ---- rotate the leg's force to the body frame (:439-444), write ----
    if (valid) {
      const bool bad = isnan(xo[0]) || isnan(xo[1]) || isnan(xo[2]);
#pragma unroll
      for (int cj = 0; cj < 3; ++cj) {
        const double g = R[cj] * xo[0] + R[3 + cj] * xo[1] + R[6 + cj] * xo[2];
        results[p].grf[3 * lg + cj] = bad ? 0.0f : (float)g;
      }
      if (lg == 0) {
        results[p].status = status_out;
        results[p].iters = iter_out;
        results[p].rho_updates = rho_out;
        results[p].pri_res = (float)pri_out;
      }
    }
  }  // next eight problems
}

}  // namespace mpcb200
