// Variant of wrench_kernel.cuh with a 2-D register tiling of the 60 x 60 core: thread (rq, q) of the tile role
// holds rows 4 rq .. 4 rq + 3 x columns 8 q .. 8 q + 7 of the swept matrix / of Y' (the matrix padded to 64 columns)
// instead of half a row.  The mat-vec omega = Y' tau then needs 8 entries of tau per thread instead of 30, and a
// sweep step 2 x 8 entries of the pivot rows instead of 2 x 30 -- the shared-memory pipe, charged per load
// instruction by width, is the kernel's bound (profiles/r02_wrench_v1_ncu_summary.txt, scripts/micro/lds_bench.cu).
// The eight partial sums of a row quad meet by a transposed reduction (four shuffles for four rows).  Everything
// else is wrench_kernel.cuh's, line for line.
#pragma once

#include "wrench_kernel.cuh"

namespace mpcb200 {

// `warm` == nullptr: cold solves.  ONE instantiation serves both, so a fresh warm slot takes the cold
// path instruction for instruction (bit-identical results: the all-four-stance states amplify even a
// different FMA contraction of two template instances into the last float32 bits).
template <int kCtas>
__global__ void __launch_bounds__(kWrThreads, kCtas)
wrench_tile_kernel(const MpcStateIn* __restrict__ states, const MpcGaitIn* __restrict__ gait,
                    MpcResult* __restrict__ results, float* __restrict__ x_all, int num, int* __restrict__ counter,
                    double* __restrict__ warm, const MpcTorqueIn* __restrict__ tin, MpcTorqueOut* __restrict__ tout,
                    const __grid_constant__ BuildParams bp, const __grid_constant__ SolveParams sp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WrenchSmem& sm = *reinterpret_cast<WrenchSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool active = lane < 30;
  const int j = 30 * warp + (active ? lane : 0);  // idle lanes shadow lane 0 (never write)
  // first lane of the leg-step, 3 (lane / 3) (30 warp is a multiple of 3; lane * 11 >> 5 = lane / 3 below 32), and the
  // component (0 fx, 1 fy, 2 fz): written so that the loop re-forms them in three instructions instead of
  // re-reading a spilled j / 3
  const int lb = active ? 3 * ((lane * 11) >> 5) : 0;
  const int comp = active ? lane - lb : 0;
  const int g = j / 3;                            // leg-step
  const int k = j / 12, jj = j - 12 * k;          // horizon step, index inside the step
  const int leg = jj / 3;
  const int r = j >> 1, h = j & 1;                // wrench role: row, half
  const int kr = k, rr = r - 6 * k;               // r / 6 = j / 12: the row's step is the variable's step
  const bool zlane = comp == 2;
  const bool kWarm = warm != nullptr;
  // element offsets used in every iteration
  const int offRowIn = 12 * kr + 6 * h;  // first rhs entry of this thread's half row
  const int offOm = 6 * k;
  const int jpin = j;
  const double mu = sp.mu, sigma = sp.sigma, alpha = sp.alpha;
  const double dt = bp.dt, inv_m = 1.0 / bp.mass;
  const double r2 = bp.Rd[jj];                    // 2 r_weights of this variable (ConvexMpc.cpp:41)
  // rows owned: A = first (or only) row, B = second row of fx / fy lanes
  const int rowA = 5 * g + (zlane ? 4 : 2 * comp), rowB = rowA + 1;

  // ---- once per CTA: alpha / beta tables of S ----
  if (tid < kH * kH) {
    const int kk = tid / kH, ll = tid - kH * kk;
    const int mxk = kk > ll ? kk : ll;
    const double half = bp.exact_discretization ? 0.5 : 0.0;
    double s = 0.0;
    for (int i = mxk; i < kH; ++i) s += ((double)(i - kk) + half) * ((double)(i - ll) + half);
    sm.al[tid] = (double)(kH - mxk) * dt * dt;
    sm.be[tid] = (dt * dt) * (dt * dt) * s;
  }

  // tile role: tid = 8 rq + q holds rows 4 rq .. 4 rq + 3 x columns 8 q .. 8 q + 7 (columns 60 .. 63 are padding)
  const bool tact = tid < 2 * kW6;            // 120 tile threads
  const int trp = tact ? (tid >> 3) : 0, tq = tid & 7;
  const int tr0 = 4 * trp;                    // first row of the tile
  double y[4][8];
  // Column blocks sit 10 doubles apart in shared memory (8 apart, the blocks 0, 2, 4, 6 of a quarter-warp's 16-byte loads
  // share banks; 10 x 8 = 80 bytes walks all eight 16-byte slots of a 128-byte line); entry r of a 60-vector is at
  // r + 2 (r / 8).
  constexpr int kTS = 10;
  auto tpad = [](int r_) { return r_ + 2 * (r_ >> 3); };
  if (tid < 4) sm.vt[kTS * 7 + 4 + tid] = 0.0;   // padding entries of tau (the padding columns of Y' are zero, 0 * NaN is not)

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.flags[3] = atomicAdd(counter, 1);
    __syncthreads();
    const int p = sm.flags[3];
    if (p >= num) break;

    // ---- K0: record load ----
    if (tid < 48) sm.st[tid] = reinterpret_cast<const float*>(states + p)[tid];
    double* const ws = kWarm ? warm + size_t(p) * kWarmStride : nullptr;
    const bool live = kWarm && ws[kWarmLive] != 0.0;
    const double rho0 = live ? ws[kWarmRho] : sp.rho;
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
    const double half = bp.exact_discretization ? 0.5 : 0.0;
    if (tid < kH) {
      // Q (A_d^(i+1) x0 - x_ref,i): A_d^m x0 = x0 + m dt A_c x0 + c2 g e_5 (A_c^2 x0 = g e_5, A_c^3 = 0)
      const int i = tid;
      const double m = (double)(i + 1);
      const double gr = -9.8;
      const double c2 = bp.exact_discretization ? 0.5 * (m * dt) * (m * dt) : 0.5 * m * (m - 1.0) * dt * dt;
      const double wx = st[kOffAngVel], wy = st[kOffAngVel + 1], wz = st[kOffAngVel + 2];
      const double vx = st[kOffLinVel], vy = st[kOffLinVel + 1], vz = st[kOffLinVel + 2];
      const double R0 = st[kOffRot + 0], R1 = st[kOffRot + 1], R2 = st[kOffRot + 2];
      const double R3 = st[kOffRot + 3], R4 = st[kOffRot + 4], R5 = st[kOffRot + 5];
      const double dx = st[kOffLinVelD], dy = st[kOffLinVelD + 1], dz = st[kOffLinVelD + 2];
      const double vwx = R0 * dx + R1 * dy + R2 * dz, vwy = R3 * dx + R4 * dy + R5 * dz;  // :470
      double xi[13], xr[13];
      xi[0] = (double)st[kOffEuler] + m * dt * (cyaw * wx + syaw * wy);
      xi[1] = (double)st[kOffEuler + 1] + m * dt * (-syaw * wx + cyaw * wy);
      xi[2] = (double)st[kOffEuler + 2] + m * dt * wz;
      xi[3] = (double)st[kOffPos] + m * dt * vx;
      xi[4] = (double)st[kOffPos + 1] + m * dt * vy;
      xi[5] = (double)st[kOffPos + 2] + m * dt * vz + c2 * gr;
      xi[6] = wx; xi[7] = wy; xi[8] = wz;
      xi[9] = vx; xi[10] = vy; xi[11] = vz + m * dt * gr;
      xi[12] = gr;
      xr[0] = (double)st[kOffEulerD];                                             // :472-488
      xr[1] = (double)st[kOffEulerD + 1];
      xr[2] = (double)st[kOffEuler + 2] + (double)st[kOffAngVelD + 2] * dt * m;
      xr[3] = (double)st[kOffPos] + vwx * dt * m;
      xr[4] = (double)st[kOffPos + 1] + vwy * dt * m;
      xr[5] = (double)st[kOffPosDz];
      xr[6] = (double)st[kOffAngVelD]; xr[7] = (double)st[kOffAngVelD + 1]; xr[8] = (double)st[kOffAngVelD + 2];
      xr[9] = vwx; xr[10] = vwy; xr[11] = 0.0; xr[12] = gr;
#pragma unroll
      for (int q = 0; q < 13; ++q) sm.Qe[i][q] = bp.Qd[q] * (xi[q] - xr[q]);
    } else if (tid >= 32 && tid < 32 + 4 * kH) {
      // B6c of (step, leg): I_w^-1 [r]x on top, I / m below (ConvexMpc.cpp:132-143); contacts
      const int t = tid - 32, stp = t >> 2, lg = t & 3;
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
      double fx = st[kOffFoot + 3 * lg], fy = st[kOffFoot + 3 * lg + 1], fz = st[kOffFoot + 3 * lg + 2];
      if (bp.foot_drift) {
        const double vx = st[kOffLinVelD], vy = st[kOffLinVelD + 1], vz = st[kOffLinVelD + 2];
        const double kd = (double)stp * dt;
        fx -= kd * (R[0] * vx + R[1] * vy + R[2] * vz);
        fy -= kd * (R[3] * vx + R[4] * vy + R[5] * vz);
        fz -= kd * (R[6] * vx + R[7] * vy + R[8] * vz);
      }
      const double sk[9] = {0.0, -fz, fy, fz, 0.0, -fx, -fy, fx, 0.0};  // Utils::skew
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          double s = 0.0;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) s += Inv[3 * i + kk] * sk[3 * kk + q];
          sm.B6[stp][i][3 * lg + q] = s;
          sm.B6[stp][3 + i][3 * lg + q] = (i == q) ? inv_m : 0.0;
        }
      int c = st[kOffContacts + lg] != 0.0f;
      if (bp.gait_aware && stp > 0) {
        // planned contact of step i from the gait counter (A1RobotControl.cpp:156-164)
        const float* gi = reinterpret_cast<const float*>(gait + p);
        const double cnt = fmod((double)gi[lg] + (double)stp * (double)gi[10] * (double)gi[4 + lg], (double)gi[8]);
        c = cnt <= (double)gi[9];
      }
      sm.contacts[4 * stp + lg] = c;
    }
    if (tid < kNP) sm.Dp[tid] = 1.0;
    __syncthreads();
    // gam_k = sum_{i >= k} F_(i-k)' Q e_i  (one thread per wrench component)
    if (tid < kW6) {
      const int kk = tid / 6, c = tid - 6 * kk;
      double s = 0.0;
      for (int i = kk; i < kH; ++i) {
        const double kp = ((double)(i - kk) + half) * dt * dt;
        const double* e = sm.Qe[i];
        double lin, rot;
        if (c == 0) { lin = e[6]; rot = cyaw * e[0] - syaw * e[1]; }
        else if (c == 1) { lin = e[7]; rot = syaw * e[0] + cyaw * e[1]; }
        else if (c == 2) { lin = e[8]; rot = e[2]; }
        else { lin = e[6 + c]; rot = e[c]; }
        s += dt * lin + kp * rot;
      }
      sm.gam[tid] = s;
    }
    __syncthreads();

    // ---- per-variable constants of the build ----
    const double top0 = sm.B6[k][0][jj], top1 = sm.B6[k][1][jj], top2 = sm.B6[k][2][jj];
    // gradient (ConvexMpc.cpp:215-217): q_j = B6c[:, j] . gam_k
    const double q0 = top0 * sm.gam[6 * k] + top1 * sm.gam[6 * k + 1] + top2 * sm.gam[6 * k + 2] +
                      inv_m * sm.gam[6 * k + 3 + comp];
    const double q_scale = live ? ws[kWarmQ + j] : q0;
    // S_(k', k) B6c[:, j] = alpha u1 + beta u2 (top three entries) and (alpha vb1 + beta vb2) m at entry 3 + comp
    double u1[3], u2[3];
    u1[0] = bp.Qd[6] * top0; u1[1] = bp.Qd[7] * top1; u1[2] = bp.Qd[8] * top2;
    u2[0] = th00 * top0 + th01 * top1; u2[1] = th01 * top0 + th11 * top1; u2[2] = th22 * top2;
    const double Qv = (comp == 0) ? bp.Qd[9] : (comp == 1) ? bp.Qd[10] : bp.Qd[11];
    const double Qp = (comp == 0) ? bp.Qd[3] : (comp == 1) ? bp.Qd[4] : bp.Qd[5];
    const double vb1 = Qv * inv_m * inv_m, vb2 = Qp * inv_m * inv_m;
    // diagonal entry of P (the only one R2 touches): P_jj = B6c_j' S_kk B6c_j + 2 r_j
    double pjj;
    {
      const double a = sm.al[kH * k + k], b = sm.be[kH * k + k];
      pjj = top0 * (a * u1[0] + b * u2[0]) + top1 * (a * u1[1] + b * u2[1]) + top2 * (a * u1[2] + b * u2[2]) +
            (a * vb1 + b * vb2) + r2;
    }
    // bounds of the owned rows (ConvexMpc.cpp:223-245); fp32 like the dense path's hand-over
    double lbA, ubA, lbB, ubB;
    if (zlane) {
      const float cflag = sm.contacts[4 * k + leg] ? 1.0f : 0.0f;
      lbA = (double)((float)bp.fz_min * cflag);
      ubA = (double)((float)bp.fz_max * cflag);
      lbB = 0.0; ubB = 0.0;
    } else {
      lbA = 0.0; ubA = (double)(float)MPC_INFTY;
      lbB = -(double)(float)MPC_INFTY; ubB = 0.0;
    }

    // ---- K3a: modified Ruiz equilibration (osqp scaling.c scale_data), nothing materialised ----
    double D = 1.0, EA = 1.0, EB = 1.0, c_run = 1.0;
    if (sp.scaling > 0) {
      double nP = max_bits(wr_colnorm(sm, bp.foot_drift != 0, k, comp, u1, u2, vb1, vb2), pjj);
      __syncthreads();  // Dp is rewritten inside the loop
      for (int it = 0; it < sp.scaling; ++it) {
        // column norms of [P; A] and row norms of A from the current D, E
        const double mE = zlane ? EA : fmax(EA, EB);
        const Leg3 e3 = leg3(mE, lb);
        const Leg3 d3 = leg3(D, lb);
        const double nA = (zlane ? fmax(mu * fmax(e3.a, e3.b), EA) : mE) * D;
        const double Dn = D * rsqrt(limit_scaling(fmax(nP, nA)));
        const double nrow = zlane ? d3.c : fmax(D, mu * d3.c);
        EA = EA * rsqrt(limit_scaling(EA * nrow));
        EB = EB * rsqrt(limit_scaling(EB * nrow));
        D = Dn;
        if (active) sm.Dp[j] = D;
        __syncthreads();
        // cost normalisation with the new D and the old c
        const double c_old = c_run;
        const double nP2 = c_old * D * max_bits(wr_colnorm(sm, bp.foot_drift != 0, k, comp, u1, u2, vb1, vb2), pjj * D);
        double part_sum = active ? nP2 : 0.0;
        double part_q = active ? fabs(c_old * D * q_scale) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          part_sum += __shfl_xor_sync(0xffffffffu, part_sum, o);
          part_q = fmax(part_q, __shfl_xor_sync(0xffffffffu, part_q, o));
        }
        double* red = sm.red + (it & 1) * (kWrWarps * 8);
        if (lane == 0) {
          red[warp * 2 + 0] = part_sum;
          red[warp * 2 + 1] = part_q;
        }
        __syncthreads();
        double s = 0.0, qn = 0.0;
#pragma unroll
        for (int w = 0; w < kWrWarps; ++w) {
          s += red[w * 2 + 0];
          qn = fmax(qn, red[w * 2 + 1]);
        }
        const double mean = s / (double)kN;
        const double ct = 1.0 / limit_scaling(fmax(mean, limit_scaling(qn)));
        c_run = c_old * ct;
        nP = nP2 * ct;
      }
    }
    const double c = c_run;
    if (tid == 0) { sm.scal[0] = c; sm.scal[1] = 1.0 / c; }
    double qb = c * D * q0;
    asm volatile("" : "+d"(qb));   // opaque: otherwise the loop re-forms c D q0 from three spilled values every iteration
    lbA *= EA; ubA *= EA; lbB *= EB; ubB *= EB;
    // constraint types of the owned rows (auxil.c set_rho_vec): -1 loose, 1 equality, 0 inequality
    auto ctype_of = [](double lo, double hi) {
      return (lo < -MPC_INFTY * 1e-4 && hi > MPC_INFTY * 1e-4) ? -1 : ((hi - lo < 1e-4) ? 1 : 0);
    };
    const int ctA = ctype_of(lbA, ubA);
    const int ctB = zlane ? 0 : ctype_of(lbB, ubB);
    // Rows in normalised form.  A row of a leg-step is  E (D_own x_own +- mu D_fz x_fz)  (fx / fy lanes) or
    // E D_fz x_fz (fz lane) = cca (x_own + s t x_fz) with cca = E D_own, t = mu D_fz / D_own (0 on the fz lane),
    // s = +1 for row A, -1 for row B.  The kernel iterates on  zh = z / cca  and  uh = y / (rho cca):
    //   zh~ = x~_own + s t x~_fz,  zh <- clip(alpha zh~ + (1 - alpha) zh + uh),  uh <- uh + (relaxed - zh),
    //   (A_'(rho z - y))_own = kapA (zhA - uhA) + kapB (zhB - uhB),   kap = rho cca^2,
    //   fz lane additionally receives  t (kapA (zhA - uhA) - kapB (zhB - uhB))  from the fx and fy lanes.
    // Same iterates as OSQP's (z, y) up to rounding, with five per-lane constants instead of ten (the ten
    // spilled out of the register file in the inner loop) and fewer multiplications.  The unscaled row
    // value E^-1 z is simply D_own zh.
    double ccaA, ccaB, tz;
    {
      const double Dz = shfl(D, lb + 2);
      ccaA = EA * D;
      ccaB = zlane ? 1.0 : EB * D;
      tz = zlane ? 0.0 : mu * Dz / D;
    }
    auto rho_of = [](int ct, double rho) { return (ct == -1) ? 1e-6 : (ct == 1) ? 1e3 * rho : rho; };
    double rvA = rho_of(ctA, rho0), rvB = rho_of(ctB, rho0);
    double kapA = rvA * ccaA * ccaA, kapB = zlane ? 0.0 : rvB * ccaB * ccaB;
    // per-thread slots (idle lanes fill theirs with the shadowed lane's finite values)
    sm.loA[tid] = lbA / ccaA; sm.hiA[tid] = ubA / ccaA;

    // iterates: x on the variable lane, zh / uh of the owned rows
    double x = 0.0, zA = 0.0, uA = 0.0, zB = 0.0, uB = 0.0;
    if (live) {
      x = ws[kWarmX + j];
      zA = ws[kWarmZ + rowA] / ccaA; uA = ws[kWarmY + rowA] / (rvA * ccaA);
      if (!zlane) { zB = ws[kWarmZ + rowB] / ccaB; uB = ws[kWarmY + rowB] / (rvB * ccaB); }
    }
    double di0 = 0.0, di1 = 0.0, di2 = 0.0;  // row `comp` of Delta_g^-1

    // rhs_j = sigma x_j - q_j + (A_'(rho z - y))_j  -> sm.va.  tau = G^ Delta^-1 rhs = M^ rhs needs only rhs, so
    // a = Delta^-1 rhs (three shuffles) is formed behind the barrier, off the critical path.
    auto publish_rhs = [&]() {
      const double eA = kapA * (zA - uA), eB = kapB * (zB - uB);
      const double own = eA + eB, sz = tz * (eA - eB);
      const double sx = shfl(sz, lb), sy = shfl(sz, lb + 1);
      const double atw = zlane ? (own + (sx + sy)) : own;
      const double rhs = sigma * x - qb + atw;
      if (active) sm.va[jpin] = rhs;
      return rhs;
    };
    auto delta_inv = [&](double rhs) {
      const Leg3 r3 = leg3(rhs, lb);
      return di0 * r3.a + di1 * r3.b + di2 * r3.c;
    };

    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    bool need_factor = true;
    double a_own = 0.0, rhs_own = 0.0;

    for (;;) {
      if (need_factor) {
        need_factor = false;
        // ---- K3b: factorisation ----
        // Delta_g = diag(c D^2 r2 + sigma) + A_g' rho A_g  (3 x 3, zero xy entry), inverse by cofactors
        {
          // own diagonal; fx / fy lanes also carry the cross term with fz and their share of the fz diagonal
          const double pd = c * D * D * r2 + sigma + kapA + kapB;
          const double pc = tz * (kapA - kapB);
          const double pz = tz * tz * (kapA + kapB);
          const Leg3 d3 = leg3(pd, lb), c3 = leg3(pc, lb), z3 = leg3(pz, lb);
          const double dxx = d3.a, dyy = d3.b, dzz = d3.c + z3.a + z3.b + z3.c, dxz = c3.a, dyz = c3.b;
          const double m00 = dyy * dzz - dyz * dyz, m11 = dxx * dzz - dxz * dxz, m22 = dxx * dyy;
          const double idet = 1.0 / (dxx * m00 - dxz * dxz * dyy);
          const double i00 = m00 * idet, i01 = dxz * dyz * idet, i02 = -dyy * dxz * idet;
          const double i11 = m11 * idet, i12 = -dxx * dyz * idet, i22 = m22 * idet;
          di0 = (comp == 0) ? i00 : (comp == 1) ? i01 : i02;
          di1 = (comp == 0) ? i01 : (comp == 1) ? i11 : i12;
          di2 = (comp == 0) ? i02 : (comp == 1) ? i12 : i22;
        }
        // M1 = G_ Delta^-1: column j = sum over the leg-step of B6c[:, j'] D_j' Delta^-1[j', j]
        {
          const Leg3 d3 = leg3(D, lb);
          const double w0 = d3.a * di0, w1 = d3.b * di1, w2 = d3.c * di2;
          const double* b6 = &sm.B6[k][0][3 * leg];
          if (active) {
#pragma unroll
            for (int cc = 0; cc < 6; ++cc)
              sm.Mh[j][cc] = b6[12 * cc] * w0 + b6[12 * cc + 1] * w1 + b6[12 * cc + 2] * w2;
          }
        }
        __syncthreads();
        // N_k = M1_k G_k' (6 x 6 per step): three entries per thread of the step
        if (active) {
#pragma unroll
          for (int e3 = 0; e3 < 3; ++e3) {
            const int e = 3 * jj + e3, c1 = e / 6, c2 = e - 6 * c1;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 12; ++i) s = fma(sm.Mh[12 * k + i][c1], sm.B6[k][c2][i] * sm.Dp[12 * k + i], s);
            sm.L[k][e] = s;
          }
        }
        __syncthreads();
        // Cholesky N_k = L L' (one thread per step; a non-positive pivot zeroes its column)
        if (active && jj == 0) {
          double Lm[6][6];
#pragma unroll
          for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) Lm[a][b] = sm.L[k][6 * a + b];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double d = Lm[cc][cc];
#pragma unroll
            for (int q = 0; q < cc; ++q) d -= Lm[cc][q] * Lm[cc][q];
            const bool ok = d > 0.0;
            const double ld = ok ? sqrt(d) : 0.0, li = ok ? 1.0 / ld : 0.0;
            Lm[cc][cc] = ld;
            sm.Linvd[k][cc] = li;
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
            for (int b = 0; b < 6; ++b) sm.L[k][6 * a + b] = (b <= a) ? Lm[a][b] : 0.0;
        }
        __syncthreads();
        // M^ = L^-1 M1 (= G^ Delta^-1 with G^ = L^-1 G_), column j by forward substitution; stored by
        // columns (x~ = a - M^' omega) and by half rows (tau = M^ rhs)
        {
          const double* Lk = sm.L[k];
          const double* li = sm.Linvd[k];
          double mcol[6];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            double smm = sm.Mh[j][cc];
#pragma unroll
            for (int q = 0; q < cc; ++q) smm -= Lk[6 * cc + q] * mcol[q];
            mcol[cc] = smm * li[cc];
          }
          if (active) {
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
              sm.Mh[j][cc] = mcol[cc];
              sm.Mr[2 * (6 * k + cc) + jj / 6][jj % 6] = mcol[cc];
            }
          }
        }
        // I + W, W = L' C L, the tile's 4 x 8 entries: W[r][6l + c] = sum_{c~ >= c} t[c~] L_l[c~][c],
        // t[c~] = c (alpha_kr,l p1[c~] + beta_kr,l p2[c~]),  p1 = D1 lam, p2 = D2 lam, lam = L_kr[:, rr]
        __syncthreads();   // (the tile role reads L of steps other warps factored)
        {
#pragma unroll
          for (int rw = 0; rw < 4; ++rw) {
            const int rt = tr0 + rw, krt = rt / 6, rrt = rt - 6 * krt;
            const double* Lk = sm.L[krt];
            double lam[6], p1[6], p2[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) lam[q] = Lk[6 * q + rrt];
            p1[0] = bp.Qd[6] * lam[0]; p1[1] = bp.Qd[7] * lam[1]; p1[2] = bp.Qd[8] * lam[2];
            p1[3] = bp.Qd[9] * lam[3]; p1[4] = bp.Qd[10] * lam[4]; p1[5] = bp.Qd[11] * lam[5];
            p2[0] = th00 * lam[0] + th01 * lam[1]; p2[1] = th01 * lam[0] + th11 * lam[1]; p2[2] = th22 * lam[2];
            p2[3] = bp.Qd[3] * lam[3]; p2[4] = bp.Qd[4] * lam[4]; p2[5] = bp.Qd[5] * lam[5];
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const int col = 8 * tq + c8;
              const bool real = col < kW6;
              const int l = real ? col / 6 : 0, cc = real ? col - 6 * l : 0;
              const double ca = c * sm.al[kH * krt + l], cb = c * sm.be[kH * krt + l];
              const double* Ll = sm.L[l];
              double sacc = (col == rt) ? 1.0 : 0.0;
#pragma unroll
              for (int q = 0; q < 6; ++q) {
                const double t = ca * p1[q] + cb * p2[q];
                if (q >= cc) sacc = fma(t, Ll[6 * q + cc], sacc);
              }
              y[rw][c8] = real ? sacc : 0.0;
            }
          }
        }
        // Symmetric sweep of A = I + W by 2 x 2 pivot blocks S = {p, p+1}, one block per barrier.  With
        // M = A_SS^-1:   A_SS <- -M,  A_rS <- A_rS M,  A_Sc <- M A_Sc,  A_rc <- A_rc - A_rS M A_Sc  (r, c not in S);
        // after all blocks the matrix is -A^-1.  A swept diagonal entry is stored PLUS ONE (1 - M_ss, then
        // updated like any other entry: it is a pivot only once), so the registers end up holding
        // I - A^-1 = Y' directly.  The block columns are written explicitly (A_rS M): producing them with
        // the uniform update cancels catastrophically when the pivots are large -- after rho has adapted
        // to its 1e-6 floor they reach 1e6 and ten digits were lost there, enough to push all-four-stance
        // states past the GRF gate (profiles/experiments/r02_wrench_sweep_cancellation.md).
        // Pivot indices are static (loop over column blocks, unrolled inside): no dynamic register index.
        // The pivot rows 2 pb, 2 pb + 1 are the rows 2 s, 2 s + 1 (s = pb & 1, static) of the eight tiles with rq == pb / 2.
        if (tact && trp == 0) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) { sm.prow[0][0][kTS * tq + cc] = y[0][cc]; sm.prow[0][1][kTS * tq + cc] = y[1][cc]; }
          if (tq == 0) {   // the thread that holds a pivot block inverts it for everybody
            const double idet = 1.0 / (y[0][0] * y[1][1] - y[0][1] * y[0][1]);
            sm.pm[0][0] = y[1][1] * idet; sm.pm[0][1] = -y[0][1] * idet; sm.pm[0][2] = y[0][0] * idet;
          }
        }
        __syncthreads();
#pragma unroll 1
        for (int qs = 0; qs < 8; ++qs) {
#pragma unroll
          for (int lc = 0; lc < 8; lc += 2) {
            const int pv = 8 * qs + lc;
            if (pv < kW6) {   // (uniform: the padding columns are never pivots)
              const int pb = pv >> 1;
              const int sp_ = (lc >> 1) & 1;           // pb & 1: which row pair of the tile holds the pivot rows
              const int buf = pb & 1;
              const double* r1 = sm.prow[buf][0];
              const double* r2 = sm.prow[buf][1];
              const double2 m1112 = *reinterpret_cast<const double2*>(sm.pm[buf]);
              const double m11 = m1112.x, m12 = m1112.y, m22 = sm.pm[buf][2];   // M = A_SS^-1, from the block's owner
              // A_r,p and A_r,p+1 of the tile's four rows, by symmetry from the pivot rows
              const double2* aAp = reinterpret_cast<const double2*>(r1 + tpad(tr0));
              const double2* aBp = reinterpret_cast<const double2*>(r2 + tpad(tr0));
              const double2 aA0 = aAp[0], aA1 = aAp[1], aB0 = aBp[0], aB1 = aBp[1];
              const double a1[4] = {aA0.x, aA0.y, aA1.x, aA1.y}, a2[4] = {aB0.x, aB0.y, aB1.x, aB1.y};
              const bool inblk = (trp == (pb >> 1));
              double g1[4], g2[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const bool prow_i = inblk && ((i >> 1) == sp_);
                const double p1v = (i & 1) ? m12 : m11, p2v = (i & 1) ? m22 : m12;
                g1[i] = prow_i ? p1v : -(a1[i] * m11 + a2[i] * m12);
                g2[i] = prow_i ? p2v : -(a1[i] * m12 + a2[i] * m22);
              }
              const double2* v1p = reinterpret_cast<const double2*>(r1 + kTS * tq);
              const double2* v2p = reinterpret_cast<const double2*>(r2 + kTS * tq);
              if (inblk) {
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const double2 v1 = v1p[hh], v2 = v2p[hh];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    if ((i >> 1) == sp_) {
                      y[i][2 * hh] = fma(g2[i], v2.x, g1[i] * v1.x);
                      y[i][2 * hh + 1] = fma(g2[i], v2.y, g1[i] * v1.y);
                    } else {
                      y[i][2 * hh] = fma(g2[i], v2.x, fma(g1[i], v1.x, y[i][2 * hh]));
                      y[i][2 * hh + 1] = fma(g2[i], v2.y, fma(g1[i], v1.y, y[i][2 * hh + 1]));
                    }
                  }
                }
              } else {
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const double2 v1 = v1p[hh], v2 = v2p[hh];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    y[i][2 * hh] = fma(g2[i], v2.x, fma(g1[i], v1.x, y[i][2 * hh]));
                    y[i][2 * hh + 1] = fma(g2[i], v2.y, fma(g1[i], v1.y, y[i][2 * hh + 1]));
                  }
                }
              }
              if (tq == qs) {
                // block columns written explicitly (see above); pivot rows: A_SS <- -M, diagonal stored plus one
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const bool prow_i = inblk && ((i >> 1) == sp_);
                  const double pa = (i & 1) ? -m12 : (1.0 - m11), pb2 = (i & 1) ? (1.0 - m22) : -m12;
                  y[i][lc] = prow_i ? pa : -g1[i];
                  y[i][lc + 1] = prow_i ? pb2 : -g2[i];
                }
              }
              if (pv + 2 < kW6 && tact && trp == ((pb + 1) >> 1)) {
                double* n0 = sm.prow[buf ^ 1][0];
                double* n1 = sm.prow[buf ^ 1][1];
                const int sn = sp_ ^ 1;             // (pb + 1) & 1
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                  n0[kTS * tq + cc] = sn ? y[2][cc] : y[0][cc];
                  n1[kTS * tq + cc] = sn ? y[3][cc] : y[1][cc];
                }
                if (tq == ((pv + 2) >> 3)) {   // this thread holds the next pivot block: columns lcn, lcn + 1 of its tile
                  const int lcn = (lc + 2) & 7;
                  const double d1 = sn ? y[2][lcn] : y[0][lcn], e12 = sn ? y[2][lcn + 1] : y[0][lcn + 1];
                  const double d2 = sn ? y[3][lcn + 1] : y[1][lcn + 1];
                  const double idet = 1.0 / (d1 * d2 - e12 * e12);
                  sm.pm[buf ^ 1][0] = d2 * idet; sm.pm[buf ^ 1][1] = -e12 * idet; sm.pm[buf ^ 1][2] = d1 * idet;
                }
              }
              __syncthreads();
            }
          }
        }
        rhs_own = publish_rhs();
        __syncthreads();
      }
      int run = until_check < until_adapt ? until_check : until_adapt;
      run = run < sp.max_iter - iter ? run : sp.max_iter - iter;
      // row B is (-INFTY E, 0] on fx / fy lanes (the lower bound can never bind) and absent on the fz lane
      // (all its coefficients are zero, z_B stays 0): one upper clamp at 0 serves both
      const double loA = sm.loA[tid], hiA = sm.hiA[tid];
#pragma unroll 1
      for (int q = 0; q < run; ++q) {
        // tau = M^ rhs (half a row per thread, halves meet by one shuffle)
        {
          const double2* gp = reinterpret_cast<const double2*>(sm.Mr[jpin]);
          const double2* ap = reinterpret_cast<const double2*>(&sm.va[offRowIn]);
          const double2 g0 = gp[0], g1 = gp[1], g2 = gp[2], a0 = ap[0], a1 = ap[1], a2 = ap[2];
          double s = g0.x * a0.x;
          double s2 = g0.y * a0.y;
          s = fma(g1.x, a1.x, s); s2 = fma(g1.y, a1.y, s2);
          s = fma(g2.x, a2.x, s); s2 = fma(g2.y, a2.y, s2);
          s += s2;
          s += shfl_xor(s, 1);
          // (active lanes: j = tid - 2 warp, so r = j / 2 = tid / 2 - warp: cheap to re-form, the loop has no register for it)
          if (active && h == 0) sm.vt[tpad((tid >> 1) - warp)] = s;
          a_own = delta_inv(rhs_own);  // not needed before x~: overlaps the barrier
        }
        __syncthreads();
        // omega = Y' tau: the tile's 4 x 8 products; the eight column blocks of a row quad meet by a transposed
        // reduction (each stage halves the rows a lane carries): lane q ends with row 2 (q & 1) + ((q >> 1) & 1)
        {
          const double2* tp = reinterpret_cast<const double2*>(&sm.vt[kTS * tq]);
          const double2 t0 = tp[0], t1 = tp[1], t2 = tp[2], t3 = tp[3];
          // one accumulator per row, the four rows interleaved: four independent chains keep the FP64 pipe fed
          const double tv[8] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
          double acc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = y[i][0] * tv[0];
#pragma unroll
          for (int cc = 1; cc < 8; ++cc)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fma(y[i][cc], tv[cc], acc[i]);
          const bool b0 = tq & 1, b1 = tq & 2;
          const double ra = (b0 ? acc[2] : acc[0]) + shfl_xor(b0 ? acc[0] : acc[2], 1);
          const double rb = (b0 ? acc[3] : acc[1]) + shfl_xor(b0 ? acc[1] : acc[3], 1);
          double v = (b1 ? rb : ra) + shfl_xor(b1 ? ra : rb, 2);
          v += shfl_xor(v, 4);
          if (tact && tq < 4) sm.vo[(tid >> 1) + 2 * (tid & 1)] = v;   // = tr0 + 2 (q & 1) + (q >> 1) for q < 4
        }
        __syncthreads();
        // x~ = a - M^' omega ; x, z, y updates of the owned rows ; next rhs
        {
          const double2* mp = reinterpret_cast<const double2*>(sm.Mh[jpin]);
          const double2* op = reinterpret_cast<const double2*>(&sm.vo[offOm]);
          const double2 m0 = mp[0], m1 = mp[1], m2 = mp[2], o0 = op[0], o1 = op[1], o2 = op[2];
          double s = fma(m0.x, o0.x, m0.y * o0.y);
          double s2 = fma(m1.x, o1.x, m1.y * o1.y);
          s = fma(m2.x, o2.x, s); s2 = fma(m2.y, o2.y, s2);
          const double xt = a_own - (s + s2);
          x = alpha * xt + (1.0 - alpha) * x;
          const double xtz = shfl(xt, lb + 2);
          const double ztA = fma(tz, xtz, xt), ztB = fma(-tz, xtz, xt);
          const double zrA = alpha * ztA + (1.0 - alpha) * zA, zrB = alpha * ztB + (1.0 - alpha) * zB;
          double znA = zrA + uA, znB = zrB + uB;
          znA = (znA < loA) ? loA : znA; znA = (znA > hiA) ? hiA : znA;
          znB = (znB > 0.0) ? 0.0 : znB;
          uA = uA + (zrA - znA);
          uB = uB + (zrB - znB);
          zA = znA; zB = znB;
          rhs_own = publish_rhs();
        }
        __syncthreads();
      }
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
      if (active) sm.xD[j] = D * x;
      __syncthreads();
      {
        // u = G (D x): half sums per (r, h)
        const double* b6 = &sm.B6[kr][rr][6 * h];
        const double* xd = &sm.xD[12 * kr + 6 * h];
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) s = fma(b6[i], xd[i], s);
        s += shfl_xor(s, 1);
        if (active && h == 0) sm.vt[r] = s;
      }
      __syncthreads();
      if (tid < kW6) {
        // w = S u = D1 (alpha u) + D2 (beta u), one thread per wrench component
        const int kk = tid / 6, cc = tid - 6 * kk;
        double sa = 0.0, sb0 = 0.0, sb1 = 0.0;
        const int o0 = (cc < 3) ? 0 : cc, o1 = (cc < 3) ? 1 : cc;
        for (int l = 0; l < kH; ++l) {
          const double a = sm.al[kH * kk + l], b = sm.be[kH * kk + l];
          sa = fma(a, sm.vt[6 * l + cc], sa);
          sb0 = fma(b, sm.vt[6 * l + o0], sb0);
          sb1 = fma(b, sm.vt[6 * l + o1], sb1);
        }
        double w;
        if (cc == 0) w = bp.Qd[6] * sa + th00 * sb0 + th01 * sb1;
        else if (cc == 1) w = bp.Qd[7] * sa + th01 * sb0 + th11 * sb1;
        else if (cc == 2) {
          double sb2 = 0.0;
          for (int l = 0; l < kH; ++l) sb2 = fma(sm.be[kH * kk + l], sm.vt[6 * l + 2], sb2);
          w = bp.Qd[8] * sa + th22 * sb2;
        } else w = bp.Qd[6 + cc] * sa + bp.Qd[cc] * sb0;
        sm.vo[tid] = w;
      }
      __syncthreads();
      double v[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) v[i] = 0.0;
      {
        const double xz = shfl(x, lb + 2);
        const double AhA = fma(tz, xz, x), AhB = fma(-tz, xz, x);      // normalised A x of the owned rows
        const double rpA = AhA - zA, rpB = zlane ? 0.0 : AhB - zB;
        if (active) {
          const double cB = zlane ? 0.0 : ccaB, DB = zlane ? 0.0 : D;
          v[0] = fmax(fabs(ccaA * rpA), fabs(cB * rpB));               // scaled primal residual
          v[1] = D * fmax(fabs(rpA), fabs(rpB));                       // unscaled: E^-1 cca = D
          v[2] = fmax(fabs(D * zA), fabs(DB * zB));
          v[3] = fmax(fabs(D * AhA), fabs(DB * AhB));
          v[4] = fmax(fabs(ccaA * zA), fabs(cB * zB));
          v[5] = fmax(fabs(ccaA * AhA), fabs(cB * AhB));
        }
        // P_ x = c D (R2 D x + G' w) ; A_' y
        const double* wv = &sm.vo[6 * k];
        const double gtw = top0 * wv[0] + top1 * wv[1] + top2 * wv[2] + inv_m * wv[3 + comp];
        const double Px = c * D * (r2 * (D * x) + gtw);
        const double fA = kapA * uA, fB = kapB * uB;                   // cca y of the owned rows
        const double own = fA + fB, sz = tz * (fA - fB);
        const double sx = shfl(sz, lb), sy = shfl(sz, lb + 1);
        const double Aty = zlane ? (own + (sx + sy)) : own;
        if (active) {
          const double Dinv = 1.0 / D;
          const double rd = Px + qb + Aty;
          v[6] = fabs(rd);
          v[7] = fabs(Dinv * rd);
          v[8] = fmax(fmax(fabs(Dinv * qb), fabs(Dinv * Aty)), fabs(Dinv * Px));
          v[9] = fmax(fmax(fabs(qb), fabs(Aty)), fabs(Px));
        }
      }
      {
        double v15[kTR];
#pragma unroll
        for (int i = 0; i < kTR; ++i) v15[i] = (i < 10) ? v[i] : 0.0;
        const double m = reduce_rows(v15, lane, MaxBitsOp());  // lanes 2i, 2i+1: quantity i
        if (!(lane & 1) && (lane >> 1) < 10) sm.red[warp * 16 + (lane >> 1)] = m;
      }
      __syncthreads();
      if (warp == 0) {
        double t = 0.0;
        if (lane < 10)
          t = fmax(fmax(sm.red[0 * 16 + lane], sm.red[1 * 16 + lane]), fmax(sm.red[2 * 16 + lane], sm.red[3 * 16 + lane]));
        double m[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) m[i] = shfl(t, i);
        if (tid == 0) {
          const double pri = m[1], dua = cinv * m[7];
          const double eps_pri = sp.eps_abs + sp.eps_rel * fmax(m[2], m[3]);
          const double eps_dua = sp.eps_abs + sp.eps_rel * cinv * m[8];
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
        const double rnA = rho_of(ctA, rho), rnB = rho_of(ctB, rho);
        uA *= rvA / rnA;   // y stays, uh = y / (rho cca) follows the new rho
        uB *= rvB / rnB;
        rvA = rnA; rvB = rnB;
        kapA = rvA * ccaA * ccaA;
        kapB = zlane ? 0.0 : rvB * ccaB * ccaB;
        need_factor = true;  // the factorisation ends by rebuilding the right-hand side with the new rho vector
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    // ---- K5: unscale, rotate the first step to the body frame, write ----
    const double xo = D * x;
    if (kWarm) {
      // keep the solver alive for the next tick -- unless this solve went wrong: a poisoned slot would
      // warm-start every later tick from NaN; it is marked dead instead (next tick = initSolver)
      const bool finite_own = isfinite(x) && isfinite(zA) && isfinite(uA) && isfinite(zB) && isfinite(uB);
      const int all_ok = __syncthreads_and(finite_own || !active);
      if (active) {
        ws[kWarmX + j] = x;
        ws[kWarmQ + j] = q0;
        ws[kWarmZ + rowA] = ccaA * zA; ws[kWarmY + rowA] = rvA * ccaA * uA;   // OSQP's scaled z, y
        if (!zlane) { ws[kWarmZ + rowB] = ccaB * zB; ws[kWarmY + rowB] = rvB * ccaB * uB; }
      }
      if (tid == 0) {
        ws[kWarmRho] = sm.scal[2];
        ws[kWarmLive] = (all_ok && isfinite(sm.scal[2])) ? 1.0 : 0.0;
      }
    }
    if (x_all != nullptr && active) x_all[size_t(p) * kN + j] = (float)xo;
    if (warp == 0) {
      // legs 0..3 of the first horizon step are variables 0..11 = lanes 0..11 of warp 0
      const Leg3 f3 = leg3(xo, lb);
      double gb = 0.0;
      const bool first = lane < 12;
      if (first) {
        const float* R = st + kOffRot;  // R' f (A1RobotControl.cpp:558-561)
        const double gv = (double)R[comp] * f3.a + (double)R[3 + comp] * f3.b + (double)R[6 + comp] * f3.c;
        const bool bad = isnan(f3.a) || isnan(f3.b) || isnan(f3.c);  // NaN guard (:559)
        gb = bad ? 0.0 : gv;
        results[p].grf[lane] = (float)gb;
      }
      if (tin != nullptr) {
        const Leg3 g3 = leg3(gb, lb);
        bool tnan = false;
        if (first) {
          const MpcTorqueIn& t = tin[p];
          const bool contact = st[kOffContacts + leg] != 0.0f;
          double tau[3];
          leg_torque(t.j_foot + 9 * leg, contact, g3.a, g3.b, g3.c, t.foot_forces_kin + 3 * leg, t.km_foot,
                     t.torques_gravity + 3 * leg, tau);
          const double tv = (comp == 0) ? tau[0] : (comp == 1) ? tau[1] : tau[2];
          tnan = isnan(tv);
          tout[p].joint_torques[lane] = tnan ? 0.0f : (float)tv;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, tnan);
        if (lane == 0) tout[p].nan_mask = (int32_t)(bal & 0xfffu);
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
