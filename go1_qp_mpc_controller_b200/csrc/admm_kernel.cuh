// K3+K4+K5: OSQP-equivalent ADMM solve of the condensed MPC QP (H = 10), one CTA per
// problem, problems pulled from an atomic counter.  Restates OSQP 0.6.x as the reference
// drives it (A1RobotControl.cpp:522-561): modified Ruiz equilibration, per-row rho,
// K = P + sigma I + A' diag(rho) A, ADMM with alpha-relaxation, residual termination
// every check_termination iterations, rho adaptation with refactorisation.
//
// Layout (history and measurements of v1..v15 in profiles/ and DESIGN.md):
//   256 threads = 8 warps; WARP w holds rows 15w..15w+14 of -K^-1 (five leg-steps), lane cg
//   holds the 15 x 4 register tile of columns {64i + 2cg, 64i + 2cg + 1 : i < 2}
//   (120 registers of the 255 per thread).
//   8 warps = 2 per SM sub-partition: the FP64-bound phases are balanced (10 warps were
//   3:3:2:2 and left a third of the sweep at the barrier, profiles/r01_v5_*).
//   The K^-1 mat-vec of an iteration uses the SYMMETRY of the tile: warp w multiplies the columns
//   R_w of its lanes' outputs by its own 15 right-hand-side entries, the eight partial vectors
//   meet in shared memory behind the ONE block barrier of the iteration, and lanes (2r, 2r+1)
//   add the eight terms of tile row r.  (Row sums by a recursive-halving shuffle transpose-
//   reduction, 16 shuffles per warp, remain for the Ruiz norms, P x and the residual maxima.)
//   Leg-step g of the warp has fx/fy/fz on lanes 6g, 6g+2, 6g+4, and lanes 6g..6g+4 own its five
//   constraint rows: the z/y update and the next right-hand side run through shuffles, x/z/y in
//   registers.
//   K^-1 comes from a BLOCKED symmetric sweep (Gauss-Jordan on the SPD matrix) over the register
//   tiles: 40 rank-3 blocks, synchronised by release/acquire flags instead of block barriers
//   (owner panel runs ahead, consumers stream; see sweep_group).
//   P arrives by one cp.async.bulk (TMA) per problem into shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_kernels.cuh"
#include "torque_map.cuh"

namespace mpcb200 {

// fine-grained cycle probes of the profiling instantiation (thread 0 of every CTA accumulates
// into registers-free shared slots; see scripts/phase_profile.py)
#define FINE_PROBE(sm, i) do { if (kProfile && threadIdx.x == 0) { const long long _t = clock64(); (sm).fine[i] += _t - (sm).fine_mark; (sm).fine_mark = _t; } } while (0)

constexpr int kSolveThreads = 256;
constexpr int kSolveWarps = kSolveThreads / 32;
constexpr int kRowGroups = 8;          // one per warp
constexpr int kTR = 15;                // tile rows = 5 leg-steps
constexpr int kTC = 4;                 // tile columns
constexpr int kLegPerWarp = 5;
// one warm-start slot (doubles): scaled x | previous unscaled q | scaled z | scaled y | rho | live
constexpr int kWarmX = 0, kWarmQ = kN, kWarmZ = 2 * kN, kWarmY = 2 * kN + kM, kWarmRho = 2 * kN + 2 * kM,
              kWarmLive = kWarmRho + 1, kWarmStride = 648;
constexpr int kPBytes = kN * kNP * 8;  // 122,880

struct SolveSmem {
  double P[kN * kNP];       // unscaled Hessian, row stride 128, pad columns zero (TMA destination)
  double rloc[kSolveWarps][16];         // right-hand side, each warp's own 15 entries (16-byte aligned, pad = 0)
  double part[2][kSolveWarps][kNP];     // K^-1 matvec: per-warp partial sums of every output, double
                                        //   buffered by iteration parity
  double xD[kNP];           // D .* x for P x (pad = 0)
  double Dp[kNP];           // D (pad = 0)
  double Vb[10][3][kNP];    // blocked sweep, ten slots: published pivot rows V' = A_S,: with A_SS - I
  double Wb[10][3][kNP];    //   W = -A_SS^-1 V' of the same block, 16 entries per row group (15 used)
  double Mb[10][8];         //   A_SS^-1: 00 01 02 11 12 22
  double G[kLegSteps * 6];  // A' diag(rho) A per leg-step: xx, xz, yy, yz, zz, (pad)
  // per-lane constants of the ADMM loop (slot [tid]); registers are kept for the K^-1 tile
  double lane_lb[kSolveThreads], lane_ub[kSolveThreads], lane_rv[kSolveThreads], lane_rinv[kSolveThreads];
  double lane_qb[kSolveThreads], lane_D[kSolveThreads], lane_Einv[kSolveThreads];
  double red[kSolveWarps * 16];
  double scal[8];           // 0:c 1:cinv 2:rho 3:ct 4:pri_res
  unsigned long long mbar;
  long long fine[16];       // profiling instantiation only
  long long fine_mark;
  int flags[8];             // 0:done 1:status 2:refactor 3:problem index 4:sweep flag wait timed out
  int pub_ready;            // blocked sweep: number of blocks published so far (monotone, release/acquire)
  int grp_done[2];          // blocked sweep, per slot set: (warps x groups that used the set) finished so far
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
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// shared-memory load that stays where it is written (the scheduler may not sink it to its first use)
__device__ __forceinline__ double lds_f64(const double* p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// Reduce fifteen per-row partials over the 32 lanes of a warp by recursive halving: at the stage
// with lane mask m a lane keeps the half of its rows whose index bit matches its own lane bit and
// hands the other half to lane ^ m, so 8 + 4 + 2 + 1 + 1 = 16 shuffles do the whole transpose-
// reduction (a naive butterfly needs 75).  Lanes 2r and 2r+1 end up with the total of tile row r
// (r < 15); lanes 30, 31 hold the padding row.  No shared memory, no warp barrier.
struct AddOp { __device__ __forceinline__ double operator()(double a, double b) const { return a + b; } };
struct MaxOp { __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); } };
// max of NON-NEGATIVE doubles through their bit patterns (ordered like the values): integer
// compares on the ALU instead of DSETP on the FP64 pipe; bit-identical result
struct MaxBitsOp {
  __device__ __forceinline__ double operator()(double a, double b) const {
    return (__double_as_longlong(a) > __double_as_longlong(b)) ? a : b;
  }
};
template <class Op>
__device__ __forceinline__ double reduce_rows(const double (&s)[kTR], int lane, Op op) {
  static_assert(kTR == 15, "recursive halving is written for 15 (padded to 16) rows");
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  double a8[8], a4[4], a2[2];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double lo = s[k], hi = (k + 8 < kTR) ? s[k + 8] : 0.0;  // 0 is neutral for both ops (max of norms)
    a8[k] = op(b4 ? hi : lo, shfl_xor(b4 ? lo : hi, 16));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) a4[k] = op(b3 ? a8[k + 4] : a8[k], shfl_xor(b3 ? a8[k] : a8[k + 4], 8));
#pragma unroll
  for (int k = 0; k < 2; ++k) a2[k] = op(b2 ? a4[k + 2] : a4[k], shfl_xor(b2 ? a4[k] : a4[k + 2], 4));
  const double a1 = op(b1 ? a2[1] : a2[0], shfl_xor(b1 ? a2[0] : a2[1], 2));
  return op(a1, shfl_xor(a1, 1));
}

// Sums over the five row lanes lbase..lbase+4 of a leg-step (lbase = 6 g inside the warp):
//   .lat on lane bl: pa@bl + pa@(bl+1);  on lane bl+2: pa@(bl+2) + pa@(bl+3)
//   .z   on lane bl+4: sum of pz over the five lanes
struct LegSums { double lat, z; };
__device__ __forceinline__ LegSums leg_reduce(double pa, double pz, int lbase) {
  LegSums r;
  r.lat = pa + __shfl_down_sync(0xffffffffu, pa, 1);
  const double t = pz + __shfl_down_sync(0xffffffffu, pz, 1);
  r.z = pz + shfl(t, lbase) + shfl(t, lbase + 2);
  return r;
}

__device__ __forceinline__ void load_cols(const double* base, int cg, double (&v)[kTC]) {
  const double2* p = reinterpret_cast<const double2*>(base + 2 * cg);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double2 t = p[32 * i];
    v[2 * i] = t.x;
    v[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ int solve_col(int cg, int jj) { return 64 * (jj >> 1) + 2 * cg + (jj & 1); }

// max_j |P_rj| D_j of the rows of this warp's row group; lanes 2r, 2r+1 get tile row r
__device__ __forceinline__ double row_norm_pass(SolveSmem& sm, int rg, int cg) {
  double dcol[kTC];
  load_cols(sm.Dp, cg, dcol);
  double m[kTR];
#pragma unroll
  for (int rr = 0; rr < kTR; ++rr) {
    double pv[kTC];
    load_cols(&sm.P[(kTR * rg + rr) * kNP], cg, pv);
    double mm = 0.0;
#pragma unroll
    for (int jj = 0; jj < kTC; ++jj) mm = MaxBitsOp()(mm, fabs(pv[jj]) * dcol[jj]);
    m[rr] = mm;
  }
  return reduce_rows(m, cg, MaxBitsOp());
}

// Blocked symmetric sweep (Gauss-Jordan on the SPD matrix), one leg-step (3 pivots
// S = {3kb, 3kb+1, 3kb+2}) per step.  With V = A_S,: (before the step), M = A_SS^-1, W = -M V:
//   A_rj <- A_rj + sum_s W[s][r] V'[s][j]   (r not in S; V' = V with A_SS - I in the S columns,
//                                            which makes the same update produce A_rS M)
//   A_Sj <- M A_Sj (j not in S),  A_SS <- -M
// (W[s][r] doubles as the column factor because A is symmetric).  Block kb lives in warp
// kb / 5, tile rows 3 (kb % 5) .. +2 (SUB is a template parameter: static register indices).
//
// Synchronisation is producer/consumer, not a block barrier per step: the owner of a block
// publishes V', M and W = -M V' for it and releases a monotone flag; every warp only acquires
// the flag of the block it is about to apply (sweep_group below).  Warps drift apart, so one
// warp's latency chain (flag, loads, publication) overlaps its sub-partition partner's FMAs and
// the step cost falls towards the FP64 pipe time.  Ten slots (two groups of five blocks) are
// recycled; a monotone "warps x groups done" counter guards the rewrite, no block barrier at all.
__device__ __forceinline__ void flag_release(int* f, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(f)), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_acquire(const int* f) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(f)) : "memory");
  return v;
}
// position of row/column r inside a published W row: 16 slots per row group (15 used)
__device__ __forceinline__ int wpos(int r) { return r + r / kTR; }

// Executed by the warp that owns block kbn once its tile rows 3 SUBN .. 3 SUBN + 2 are up to date.
template <int SUBN>
__device__ __forceinline__ void publish_block(SolveSmem& sm, const double (&a)[kTR][kTC], int kbn, int cg, int slot,
                                              int flag_value) {
  const int c0 = 3 * kbn;
  double vp[3][kTC];  // V' entries of this lane
#pragma unroll
  for (int s3 = 0; s3 < 3; ++s3) {
    double2* dst = reinterpret_cast<double2*>(&sm.Vb[slot][s3][2 * cg]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int col = 64 * i + 2 * cg;
      double v0 = a[3 * SUBN + s3][2 * i], v1 = a[3 * SUBN + s3][2 * i + 1];
      if (col == c0 + s3) v0 -= 1.0;
      if (col + 1 == c0 + s3) v1 -= 1.0;
      vp[s3][2 * i] = v0;
      vp[s3][2 * i + 1] = v1;
      dst[32 * i] = make_double2(v0, v1);
    }
  }
  __syncwarp();
  // A_SS (identity added back) and its symmetric 3x3 cofactor inverse, redundantly on every lane
  const double(*V)[kNP] = sm.Vb[slot];
  const double m00 = V[0][c0] + 1.0, m01 = V[0][c0 + 1], m02 = V[0][c0 + 2];
  const double m11 = V[1][c0 + 1] + 1.0, m12 = V[1][c0 + 2];
  const double m22 = V[2][c0 + 2] + 1.0;
  const double k00 = m11 * m22 - m12 * m12, k01 = m02 * m12 - m01 * m22, k02 = m01 * m12 - m02 * m11;
  const double id = __drcp_rn(m00 * k00 + m01 * k01 + m02 * k02);
  const double i00 = k00 * id, i01 = k01 * id, i02 = k02 * id;
  const double i11 = (m00 * m22 - m02 * m02) * id, i12 = (m01 * m02 - m00 * m12) * id;
  const double i22 = (m00 * m11 - m01 * m01) * id;
  // W = -M V' for this lane's four columns
#pragma unroll
  for (int jj = 0; jj < kTC; ++jj) {
    const int col = solve_col(cg, jj);
    if (col < kN) {  // pad columns have no row
      const int wp = wpos(col);
      const double x0 = vp[0][jj], x1 = vp[1][jj], x2 = vp[2][jj];
      sm.Wb[slot][0][wp] = -(i00 * x0 + i01 * x1 + i02 * x2);
      sm.Wb[slot][1][wp] = -(i01 * x0 + i11 * x1 + i12 * x2);
      sm.Wb[slot][2][wp] = -(i02 * x0 + i12 * x1 + i22 * x2);
    }
  }
  if (cg == 0) {
    double* m = sm.Mb[slot];
    m[0] = i00; m[1] = i01; m[2] = i02; m[3] = i11; m[4] = i12; m[5] = i22;
  }
  __syncwarp();
  if (cg == 0) flag_release(&sm.pub_ready, flag_value);
}

// rows [R0, R0 + NR) of the rank-3 update with the block published in `slot`
template <int R0, int NR>
__device__ __forceinline__ void update_rows(SolveSmem& sm, double (&a)[kTR][kTC], int slot, int rg, int cg) {
  if (NR <= 0) return;
#pragma unroll
  for (int s3 = 0; s3 < 3; ++s3) {
    double v[kTC], w[16];
    load_cols(sm.Vb[slot][s3], cg, v);
    {
      const double2* wp = reinterpret_cast<const double2*>(&sm.Wb[slot][s3][16 * rg]);
#pragma unroll
      for (int h = R0 / 2; h < (R0 + NR + 1) / 2; ++h) {
        const double2 t = wp[h];
        w[2 * h] = t.x;
        w[2 * h + 1] = t.y;
      }
    }
#pragma unroll
    for (int rr = R0; rr < R0 + NR; ++rr) {
#pragma unroll
      for (int jj = 0; jj < kTC; ++jj) a[rr][jj] = fma(w[rr], v[jj], a[rr][jj]);
    }
  }
}

// pivot rows of block kb (tile rows 3 SUB ..): A_Sj <- M A_Sj (j not in S), A_SS <- -M
template <int SUB>
__device__ __forceinline__ void pivot_rows(SolveSmem& sm, double (&a)[kTR][kTC], int kb, int slot, int cg) {
  const int c0 = 3 * kb;
  const double* m = sm.Mb[slot];
  const double i00 = m[0], i01 = m[1], i02 = m[2], i11 = m[3], i12 = m[4], i22 = m[5];
#pragma unroll
  for (int jj = 0; jj < kTC; ++jj) {
    const double x0 = a[3 * SUB][jj], x1 = a[3 * SUB + 1][jj], x2 = a[3 * SUB + 2][jj];
    const int t = solve_col(cg, jj) - c0;  // position inside S, if any
    a[3 * SUB][jj] = (t == 0) ? -i00 : (t == 1) ? -i01 : (t == 2) ? -i02 : (i00 * x0 + i01 * x1 + i02 * x2);
    a[3 * SUB + 1][jj] = (t == 0) ? -i01 : (t == 1) ? -i11 : (t == 2) ? -i12 : (i01 * x0 + i11 * x1 + i12 * x2);
    a[3 * SUB + 2][jj] = (t == 0) ? -i02 : (t == 1) ? -i12 : (t == 2) ? -i22 : (i02 * x0 + i12 * x1 + i22 * x2);
  }
}

// bounded spins: a publication is at most a few thousand cycles away; a lost one must not hang the GPU
__device__ __forceinline__ void wait_flag(SolveSmem& sm, const int* f, int want) {
  int spins = 0;
  while (flag_acquire(f) < want) {
    if (++spins > (1 << 20)) { sm.flags[4] = 1; break; }
  }
}

// One group of the sweep = the five leg-steps whose pivot rows live in warp kp.
//   warp kp (owner): PANEL first -- with block j (already published) update only the rows of its
//     later blocks j+1..4 and publish block j+1 at once, so publications run ahead of the
//     consumers; then the DEFERRED part: pivot transform of block j and the update of the rows of
//     the earlier blocks, j = 0..4.  (Row updates with different blocks are additive and commute;
//     each row only needs its own pivot transform in sequence, which this order keeps.)
//   other warps: apply the five blocks as their flags come in; warp kp + 1 updates its rows 0..2
//     first on the last block and publishes block 0 of the next group.
// A slot set is rewritten two groups later: the first write waits until every warp has reported
// the set's previous group done.  The counters are PER SET: nobody can finish the set's next
// group before its first block is published, so the count is exact.  (One counter over all groups
// is not: seven fast warps finishing two groups outvote a slow owner still reading its slots --
// seen on the first problem of a CTA, when the owner's code is cold in the instruction cache.)
template <bool kProfile>
__device__ __forceinline__ void sweep_group(SolveSmem& sm, double (&a)[kTR][kTC], int kp, int rg, int cg, int base,
                                            int done_base) {
  const int set = (kp & 1) * kLegPerWarp;
  const int kb0 = kLegPerWarp * kp;
  if (rg == kp) {
    wait_flag(sm, &sm.pub_ready, base + kb0 + 1);
    FINE_PROBE(sm, 0);  // flag wait
    // per block j: its successor's three rows first, publish, then the rows of the later blocks
    update_rows<3, 3>(sm, a, set + 0, rg, cg);
    publish_block<1>(sm, a, kb0 + 1, cg, set + 1, base + kb0 + 2);
    update_rows<6, 9>(sm, a, set + 0, rg, cg);
    update_rows<6, 3>(sm, a, set + 1, rg, cg);
    publish_block<2>(sm, a, kb0 + 2, cg, set + 2, base + kb0 + 3);
    update_rows<9, 6>(sm, a, set + 1, rg, cg);
    update_rows<9, 3>(sm, a, set + 2, rg, cg);
    publish_block<3>(sm, a, kb0 + 3, cg, set + 3, base + kb0 + 4);
    update_rows<12, 3>(sm, a, set + 2, rg, cg);
    update_rows<12, 3>(sm, a, set + 3, rg, cg);
    publish_block<4>(sm, a, kb0 + 4, cg, set + 4, base + kb0 + 5);
    FINE_PROBE(sm, 5);  // panel
    pivot_rows<0>(sm, a, kb0 + 0, set + 0, cg);
    pivot_rows<1>(sm, a, kb0 + 1, set + 1, cg);
    update_rows<0, 3>(sm, a, set + 1, rg, cg);
    pivot_rows<2>(sm, a, kb0 + 2, set + 2, cg);
    update_rows<0, 6>(sm, a, set + 2, rg, cg);
    pivot_rows<3>(sm, a, kb0 + 3, set + 3, cg);
    update_rows<0, 9>(sm, a, set + 3, rg, cg);
    pivot_rows<4>(sm, a, kb0 + 4, set + 4, cg);
    update_rows<0, 12>(sm, a, set + 4, rg, cg);
    FINE_PROBE(sm, 3);  // deferred part
  } else {
    // The owner's sub-partition partner (warp kp ^ 4 shares its FP64 pipe) stays out of the way until
    // the panel -- the serial chain every other warp waits for -- has published its last block, and
    // applies the five blocks afterwards with the pipe to itself.
    if (rg == (kp ^ 4)) wait_flag(sm, &sm.pub_ready, base + kb0 + kLegPerWarp);
#pragma unroll 1
    for (int sub = 0; sub < kLegPerWarp - 1; ++sub) {
      wait_flag(sm, &sm.pub_ready, base + kb0 + sub + 1);
      FINE_PROBE(sm, 0);  // flag wait
      update_rows<0, kTR>(sm, a, set + sub, rg, cg);
      FINE_PROBE(sm, 4);  // rank-3 update
    }
    wait_flag(sm, &sm.pub_ready, base + kb0 + kLegPerWarp);
    FINE_PROBE(sm, 0);
    if (rg == kp + 1) {
      update_rows<0, 3>(sm, a, set + 4, rg, cg);
      // first write into the other slot set in its new life: everyone must be done with group kp - 1
      if (kp >= 1) wait_flag(sm, &sm.grp_done[(kp + 1) & 1], done_base + kSolveWarps * ((kp + 1) >> 1));
      publish_block<0>(sm, a, kb0 + kLegPerWarp, cg, ((kp + 1) & 1) * kLegPerWarp, base + kb0 + kLegPerWarp + 1);
      update_rows<3, 12>(sm, a, set + 4, rg, cg);
    } else {
      update_rows<0, kTR>(sm, a, set + 4, rg, cg);
    }
    FINE_PROBE(sm, 4);
  }
  // this warp no longer reads the slots of group kp
  __syncwarp();
  if (cg == 0)
    asm volatile("red.release.cta.shared::cta.add.s32 [%0], 1;" ::"r"(smem_u32(&sm.grp_done[kp & 1])) : "memory");
}

// Build K = c D P D + sigma I + A' diag(rho) A into the register tiles, then overwrite it
// with -K^-1 by the blocked symmetric sweep (40 rank-3 steps).  `base` counts the blocks this CTA
// has published so far (the flag is monotone over the CTA's life).
template <bool kProfile>
__device__ __forceinline__ void factor_inverse(SolveSmem& sm, double (&a)[kTR][kTC], int rg, int cg, double sigma,
                                               int& base, int& done_base) {
  {
    const double c = sm.scal[0];
    double dcol[kTC];
    load_cols(sm.Dp, cg, dcol);
#pragma unroll
    for (int rr = 0; rr < kTR; ++rr) {
      const int row = kTR * rg + rr;
      const int ls = kLegPerWarp * rg + rr / 3, rc = rr % 3;  // leg-step and component of this row (static rc)
      // G = [[xx, 0, xz], [0, yy, yz], [xz, yz, zz]] of the row's leg-step: row rc of it, loaded
      // unconditionally so that the column test below is three selects and no branch
      const double* g = &sm.G[ls * 6];
      const double gr0 = (rc == 0) ? g[0] : (rc == 1) ? 0.0 : g[1];
      const double gr1 = (rc == 0) ? 0.0 : (rc == 1) ? g[2] : g[3];
      const double gr2 = (rc == 0) ? g[1] : (rc == 1) ? g[3] : g[4];
      const double cDr = c * sm.Dp[row];
      double pv[kTC];
      load_cols(&sm.P[row * kNP], cg, pv);
#pragma unroll
      for (int jj = 0; jj < kTC; ++jj) {
        const int col = solve_col(cg, jj);
        const int cc = col - 3 * ls;
        // same order of additions as the branching form (adding 0.0 is exact)
        double val = cDr * pv[jj] * dcol[jj];
        val += (col == row) ? sigma : 0.0;
        val += (cc == 0) ? gr0 : (cc == 1) ? gr1 : (cc == 2) ? gr2 : 0.0;
        a[rr][jj] = val;
      }
    }
  }
  // callers arrive here behind a block barrier: nobody reads the slots of an earlier factorisation
  if (rg == 0) publish_block<0>(sm, a, 0, cg, 0, base + 1);
  if (kProfile && threadIdx.x == 0) sm.fine_mark = clock64();
  for (int kp = 0; kp < kRowGroups; ++kp) sweep_group<kProfile>(sm, a, kp, rg, cg, base, done_base);
  base += kLegSteps;
  done_base += kSolveWarps * (kRowGroups / 2);  // four groups per slot set and factorisation
}

template <bool kProfile, bool kWarm>
__global__ void __launch_bounds__(kSolveThreads, 1)
admm_solve_kernel(const double* __restrict__ P_all, const double* __restrict__ q_all,
                  const float* __restrict__ l_all, const float* __restrict__ u_all,
                  const MpcStateIn* __restrict__ states, MpcResult* __restrict__ results,
                  float* __restrict__ x_all, int num, int* __restrict__ counter,
                  long long* __restrict__ phase_clk, double* __restrict__ warm,
                  const MpcTorqueIn* __restrict__ tin, MpcTorqueOut* __restrict__ tout,
                  const __grid_constant__ SolveParams sp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SolveSmem& sm = *reinterpret_cast<SolveSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = warp, cg = lane;  // row group = warp, column group = lane
  // lane roles inside the warp (five leg-steps: lanes 6g .. 6g+4, g < 5)
  const int lg = lane / 6;
  const int pos = lane - 6 * lg;
  const bool active = lane < 30;
  const bool rown = active && pos <= 4;                            // owns constraint row `pos` of its leg-step
  const bool vown = active && (pos == 0 || pos == 2 || pos == 4);  // owns variable component pos / 2
  const int vc = pos >> 1;
  const int ls = kLegPerWarp * rg + (active ? lg : 0);             // leg-step
  const int vj = 3 * ls + (vown ? vc : 0);
  const int ri = 5 * ls + (rown ? pos : 0);
  const int lbase = active ? 6 * lg : 24;
  const int latsrc = lbase + ((pos < 2) ? 0 : 2), zsrc = lbase + 4;
  const double mu = sp.mu;
  const double sigma = sp.sigma, alpha = sp.alpha;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;
  int sweep_base = 0, sweep_done = 0;  // blocks published / (warps x groups) finished by this CTA so far
  if (tid == 0) { sm.pub_ready = 0; sm.grp_done[0] = 0; sm.grp_done[1] = 0; }
  if (kProfile && tid == 0) {
    for (int i = 0; i < 16; ++i) sm.fine[i] = 0;
    sm.fine_mark = clock64();
  }
  double a[kTR][kTC];  // register tile of -K^-1
  // per-phase cycle counters of thread 0: compiled in only for the kProfile instantiation
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

    // kWarm: slot p of `warm` is the solver this robot keeps alive between control ticks
    // (A1RobotControl.cpp:522-538).  A live slot makes this an update + warm solve: scale_data
    // still sees the PREVIOUS tick's gradient (osqp_update_P runs before osqp_update_lin_cost),
    // and x, z, y (in the old scaled coordinates) and rho carry over.
    double* const ws = kWarm ? warm + size_t(p) * kWarmStride : nullptr;
    const bool live = kWarm && ws[kWarmLive] != 0.0;
    const double rho0 = live ? ws[kWarmRho] : sp.rho;
    if (tid < kNP) {
      sm.Dp[tid] = (tid < kN) ? 1.0 : 0.0;
      sm.rloc[tid >> 4][tid & 15] = 0.0;
      sm.xD[tid] = 0.0;
    }
    if (tid == 0) {
      sm.scal[0] = 1.0;
      sm.scal[2] = rho0;
      sm.flags[0] = 0;
      sm.flags[4] = 0;
      sm.flags[1] = MPC_STATUS_UNSOLVED;
    }
    const double q0 = vown ? q_all[size_t(p) * kN + vj] : 0.0;
    const double q_scale = (live && vown) ? ws[kWarmQ + vj] : q0;
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
    double c_run = 1.0;       // the cost scaling c, carried by every thread
    if (sp.scaling > 0) {
      double nP = row_norm_pass(sm, rg, cg);  // c = 1, D = 1 (valid on the variable lanes)
      __syncthreads();                        // Dp is rewritten inside the loop
      for (int it = 0; it < sp.scaling; ++it) {
        // column norm of [P; A] for the owned variable, row norm of A for the owned row
        const double E0 = shfl(E, lbase), E1 = shfl(E, lbase + 1), E2 = shfl(E, lbase + 2),
                     E3 = shfl(E, lbase + 3), E4 = shfl(E, lbase + 4);
        const double Dx = shfl(D, lbase), Dy = shfl(D, lbase + 2), Dz = shfl(D, lbase + 4);
        double Dn = D, En = E;
        if (vown) {
          double nA;
          if (vc == 0) nA = fmax(E0, E1);
          else if (vc == 1) nA = fmax(E2, E3);
          else nA = fmax(mu * fmax(fmax(E0, E1), fmax(E2, E3)), E4);
          nA *= D;
          Dn = D * rsqrt(limit_scaling(fmax(nP, nA)));
          sm.Dp[vj] = Dn;
        }
        if (rown) {
          const double nrow = (pos == 4) ? Dz : fmax((pos < 2) ? Dx : Dy, mu * Dz);
          En = E * rsqrt(limit_scaling(E * nrow));
        }
        D = Dn;
        E = En;
        __syncthreads();
        // cost normalisation with the new D and the old c; every thread combines the eight warp
        // partials itself (same order, same bits): one block barrier, no serial section
        const double c_old = c_run;
        const double nP2 = c_old * D * row_norm_pass(sm, rg, cg);
        double part_sum = vown ? nP2 : 0.0;
        double part_q = vown ? fabs(c_old * D * q_scale) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          part_sum += __shfl_xor_sync(0xffffffffu, part_sum, o);
          part_q = fmax(part_q, __shfl_xor_sync(0xffffffffu, part_q, o));
        }
        double* red = sm.red + (it & 1) * (kSolveWarps * 8);  // double buffered by pass parity
        if (lane == 0) {
          red[warp * 2 + 0] = part_sum;
          red[warp * 2 + 1] = part_q;
        }
        __syncthreads();
        double s = 0.0, qn = 0.0;
#pragma unroll
        for (int w = 0; w < kSolveWarps; ++w) {
          s += red[w * 2 + 0];
          qn = fmax(qn, red[w * 2 + 1]);
        }
        const double mean = s / (double)kN;
        const double ct = 1.0 / limit_scaling(fmax(mean, limit_scaling(qn)));
        c_run = c_old * ct;
        nP = nP2 * ct;
      }
      if (tid == 0) sm.scal[0] = c_run;
    }
    // ---- scaled data on the owning lanes (constants go to per-lane shared-memory slots) ----
    const double c = c_run;
    const double qb0 = c * D * q0;
    lb *= E;
    ub *= E;
    int ctype = 0;
    if (lb < -MPC_INFTY * 1e-4 && ub > MPC_INFTY * 1e-4) ctype = -1;
    else if (ub - lb < 1e-4) ctype = 1;
    const double rv0 = (ctype == -1) ? 1e-6 : (ctype == 1) ? 1e3 * rho0 : rho0;
    // scaled constraint coefficients of the owned row: z~_i = cca * x~_lat + ccz * x~_z
    double cca, ccz;
    {
      const double Dlat = shfl(D, latsrc), Dz = shfl(D, zsrc);
      cca = (pos == 4) ? 0.0 : E * Dlat;
      ccz = (pos == 4) ? E * Dz : ((pos & 1) ? -mu : mu) * E * Dz;
      if (!rown) { cca = 0.0; ccz = 0.0; }
    }
    sm.lane_lb[tid] = lb;
    sm.lane_ub[tid] = ub;
    sm.lane_rv[tid] = rv0;
    sm.lane_rinv[tid] = 1.0 / rv0;
    sm.lane_qb[tid] = qb0;
    sm.lane_D[tid] = D;
    sm.lane_Einv[tid] = 1.0 / E;
    if (tid == 0) sm.scal[1] = 1.0 / c;
    auto build_G = [&]() {
      // A' diag(rho) A of this lane's leg-step from its five row lanes
      const double r0 = rown ? sm.lane_rv[tid] : 0.0;
      const LegSums s1 = leg_reduce(r0 * cca * cca, r0 * ccz * ccz, lbase);  // xx|yy on lanes 0|2, zz on lane 4
      const LegSums s2 = leg_reduce(r0 * cca * ccz, 0.0, lbase);             // xz|yz on lanes 0|2
      if (active && pos == 0) { sm.G[ls * 6 + 0] = s1.lat; sm.G[ls * 6 + 1] = s2.lat; }
      if (active && pos == 2) { sm.G[ls * 6 + 2] = s1.lat; sm.G[ls * 6 + 3] = s2.lat; }
      if (active && pos == 4) sm.G[ls * 6 + 4] = s1.z;
    };
    build_G();
    double x = 0.0, z = 0.0, y = 0.0;  // x on variable lanes; z, y on row lanes
    if (live) {
      if (vown) x = ws[kWarmX + vj];
      if (rown) { z = ws[kWarmZ + ri]; y = ws[kWarmY + ri]; }
      const double w = rown ? (rv0 * z - y) : 0.0;
      const LegSums s = leg_reduce(cca * w, ccz * w, lbase);
      if (vown) sm.rloc[rg][vj - kTR * rg] = sigma * x - qb0 + ((vc == 2) ? s.z : s.lat);
    } else {
      if (vown) sm.rloc[rg][vj - kTR * rg] = -qb0;  // rhs of iteration 1: x = z = y = 0
    }
    __syncthreads();

    PHASE_MARK(0);
    // ---- K4: ADMM iterations (osqp.c osqp_solve) ----
    int iter = 0, rho_updates = 0, status = MPC_STATUS_UNSOLVED;
    double pri_res_out = 0.0;
    // countdowns instead of iter % interval (runtime divisors cost ~50 instructions per iteration)
    int until_check = sp.check_termination > 0 ? sp.check_termination : 0x7fffffff;
    int until_adapt = (sp.adaptive_rho && sp.adaptive_rho_interval > 0) ? sp.adaptive_rho_interval : 0x7fffffff;
    bool need_factor = true;
    int par = 0;  // buffer parity of the partial sums
    // Outer loop = one stretch of iterations up to the next event (termination check, rho adaptation,
    // iteration limit); the inner loop runs on a plain counter, so an ordinary iteration carries no
    // bookkeeping of the three countdowns.
    for (;;) {
      if (need_factor) {
        // ---- K3b: factor (explicit inverse in registers); ONE call site keeps the code small ----
        need_factor = false;
        factor_inverse<kProfile>(sm, a, rg, cg, sigma, sweep_base, sweep_done);
        PHASE_MARK(1);
      }
      int run = until_check < until_adapt ? until_check : until_adapt;
      run = run < sp.max_iter - iter ? run : sp.max_iter - iter;
#pragma unroll 1
      for (int q = 0; q < run; ++q) {
      par ^= 1;
      if (kProfile && tid == 0) sm.fine_mark = clock64();
      // x~ = K^-1 rhs.  K^-1 is symmetric, so the 15 x 4 tile of rows R_w is also the tile of COLUMNS R_w
      // of this lane's four outputs: the warp multiplies by ITS OWN 15 right-hand-side entries (which
      // it wrote itself: a warp barrier, no block barrier, no operand exchange) and obtains one partial
      // sum of all 128 outputs, four per lane, with no reduction across lanes.  The eight partial
      // vectors meet in shared memory behind the iteration's one block barrier; lanes 2r, 2r+1 then
      // add the eight terms of tile row r.  (This replaces a 16-shuffle transpose-reduction per warp.)
      __syncwarp();
      double xt, lc_rv, lc_rinv, lc_lo, lc_hi, lc_qb;
      {
        double rv15[16];
        const double2* rp = reinterpret_cast<const double2*>(sm.rloc[rg]);
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const double2 t = rp[h];
          rv15[2 * h] = t.x;
          rv15[2 * h + 1] = t.y;
        }
        double s[kTC], u[kTC];
#pragma unroll
        for (int jj = 0; jj < kTC; ++jj) {
          s[jj] = a[0][jj] * rv15[0];
          u[jj] = a[1][jj] * rv15[1];
        }
#pragma unroll
        for (int rr = 2; rr + 1 < kTR; rr += 2) {
#pragma unroll
          for (int jj = 0; jj < kTC; ++jj) {
            s[jj] = fma(a[rr][jj], rv15[rr], s[jj]);
            u[jj] = fma(a[rr + 1][jj], rv15[rr + 1], u[jj]);
          }
        }
#pragma unroll
        for (int jj = 0; jj < kTC; ++jj) s[jj] = fma(a[kTR - 1][jj], rv15[kTR - 1], s[jj]) + u[jj];
        double2* outp = reinterpret_cast<double2*>(&sm.part[par][rg][2 * cg]);
        outp[0] = make_double2(s[0], s[1]);
        outp[32] = make_double2(s[2], s[3]);
        if (kProfile && s[0] == 1.2345e300) sm.fine[15] += 1;
        FINE_PROBE(sm, 9);  // iteration: rhs load + 60 FMA + partial store
        __syncthreads();
        FINE_PROBE(sm, 8);  // iteration: barrier
        // the five per-lane constants of the z / y / rhs chain are requested here, behind the barrier
        // and ahead of the gather (volatile asm keeps them from sinking to their first use)
        lc_rv = lds_f64(&sm.lane_rv[tid]);
        lc_rinv = lds_f64(&sm.lane_rinv[tid]);
        lc_lo = lds_f64(&sm.lane_lb[tid]);
        lc_hi = lds_f64(&sm.lane_ub[tid]);
        lc_qb = lds_f64(&sm.lane_qb[tid]);
        const int r = (lane >> 1) < kTR ? (lane >> 1) : 0;
        const double* pp = &sm.part[par][0][kTR * rg + r];
        double p8[kSolveWarps];
#pragma unroll
        for (int w = 0; w < kSolveWarps; ++w) p8[w] = pp[w * kNP];
        xt = -(((p8[0] + p8[1]) + (p8[2] + p8[3])) + ((p8[4] + p8[5]) + (p8[6] + p8[7])));
        if (kProfile && xt == 1.2345e300) sm.fine[15] += 1;
        FINE_PROBE(sm, 10);  // iteration: gather of the eight partial sums
      }
      // x <- alpha x~ + (1 - alpha) x
      x = alpha * xt + (1.0 - alpha) * x;
      const double sxq = sigma * x - lc_qb;  // off the z / y chain: only the leg sums are added at its end
      // z~ = A x~ ; z, y update on the row lanes
      const double rvv = lc_rv;
      {
        const double xt_lat = shfl(xt, latsrc), xt_z = shfl(xt, zsrc);
        const double zt = cca * xt_lat + ccz * xt_z;
        const double zr = alpha * zt + (1.0 - alpha) * z;
        double zn = zr + lc_rinv * y;
        {
          // box projection; plain compare-selects (double fmin/fmax cost ~25 cycles each here)
          const double lo = lc_lo, hi = lc_hi;
          zn = (zn < lo) ? lo : zn;
          zn = (zn > hi) ? hi : zn;
        }
        y = y + rvv * (zr - zn);
        z = zn;
      }
      if (kProfile && z == 1.2345e300) sm.fine[15] += 1;
      FINE_PROBE(sm, 11);  // iteration: z / y update
      // next rhs = sigma x - q + A'(rho z - y)
      {
        const double w = rown ? (rvv * z - y) : 0.0;
        const LegSums s = leg_reduce(cca * w, ccz * w, lbase);
        if (vown) sm.rloc[rg][vj - kTR * rg] = sxq + ((vc == 2) ? s.z : s.lat);
      }
      FINE_PROBE(sm, 12);  // iteration: next rhs
      }  // stretch of ordinary iterations
      iter += run;
      until_check -= run;
      until_adapt -= run;
      const bool can_check = (until_check == 0);
      const bool can_adapt = (until_adapt == 0);
      if (can_check) until_check = sp.check_termination;
      if (can_adapt) until_adapt = sp.adaptive_rho_interval;
      const bool last = (iter == sp.max_iter);
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
        double xv[kTC];
        load_cols(sm.xD, cg, xv);
        double s[kTR];
#pragma unroll
        for (int rr = 0; rr < kTR; ++rr) {
          double pv[kTC];
          load_cols(&sm.P[(kTR * rg + rr) * kNP], cg, pv);
          s[rr] = fma(pv[0], xv[0], pv[1] * xv[1]) + fma(pv[2], xv[2], pv[3] * xv[3]);
        }
        const double sr = reduce_rows(s, lane, AddOp());
        const double yy = rown ? y : 0.0;
        const LegSums ay = leg_reduce(cca * yy, ccz * yy, lbase);
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
      {
        // ten non-negative maxima per warp in one transpose-reduction (16 shuffles instead of 50)
        double v15[kTR];
#pragma unroll
        for (int i = 0; i < kTR; ++i) v15[i] = (i < 10) ? v[i] : 0.0;
        const double m = reduce_rows(v15, lane, MaxBitsOp());  // lanes 2i, 2i+1: quantity i
        if (!(lane & 1) && (lane >> 1) < 10) sm.red[warp * 16 + (lane >> 1)] = m;
      }
      __syncthreads();
      if (warp == 0) {
        // lane i < 10 combines quantity i over the 8 warps (independent loads, 3-level tree);
        // thread 0 then gathers the ten maxima by shuffle and decides
        double t = 0.0;
        if (lane < 10) {
          const double t0 = fmax(sm.red[0 * 16 + lane], sm.red[1 * 16 + lane]);
          const double t1 = fmax(sm.red[2 * 16 + lane], sm.red[3 * 16 + lane]);
          const double t2 = fmax(sm.red[4 * 16 + lane], sm.red[5 * 16 + lane]);
          const double t3 = fmax(sm.red[6 * 16 + lane], sm.red[7 * 16 + lane]);
          t = fmax(fmax(t0, t1), fmax(t2, t3));
        }
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
          const LegSums s = leg_reduce(cca * w, ccz * w, lbase);
          if (vown) sm.rloc[rg][vj - kTR * rg] = sigma * x - sm.lane_qb[tid] + ((vc == 2) ? s.z : s.lat);
        }
        build_G();
        __syncthreads();
        need_factor = true;
      }
    }
    if (iter > sp.max_iter) iter = sp.max_iter;

    if (kWarm) {
      // keep the solver alive for the next tick -- unless this solve went wrong (non-finite iterates
      // from one bad record, or the sweep's flag wait timed out): a poisoned slot would warm-start
      // every later tick of this robot from NaN, so it is left dead and the next tick is an initSolver
      const bool own_ok = (!vown || isfinite(x)) && (!rown || (isfinite(z) && isfinite(y)));
      const int all_ok = __syncthreads_and(own_ok);
      if (vown) { ws[kWarmX + vj] = x; ws[kWarmQ + vj] = q0; }
      if (rown) { ws[kWarmZ + ri] = z; ws[kWarmY + ri] = y; }
      if (tid == 0) {
        ws[kWarmRho] = sm.scal[2];
        ws[kWarmLive] = (all_ok && !sm.flags[4] && isfinite(sm.scal[2])) ? 1.0 : 0.0;
      }
    }
    // ---- K5: unscale, rotate the first step to the body frame, write ----
    const double xo = sm.lane_D[tid] * x;
    if (x_all != nullptr && vown) x_all[size_t(p) * kN + vj] = (float)xo;
    if (rg == 0) {
      // legs 0..3 of the first horizon step are leg-steps 0..3 = lanes 0..23 of warp 0;
      // f = (x, y, z) on lanes lbase, lbase + 2, lbase + 4
      const double f0 = shfl(xo, lbase), f1 = shfl(xo, lbase + 2), f2 = shfl(xo, lbase + 4);
      double gb = 0.0;  // body-frame component vc of this leg, NaN-guarded
      if (vown && lg < 4) {
        double g;
        if (states != nullptr) {
          // R' f (A1RobotControl.cpp:558-561)
          const float* R = reinterpret_cast<const float*>(states + p) + kOffRot;
          g = (double)R[vc] * f0 + (double)R[3 + vc] * f1 + (double)R[6 + vc] * f2;
        } else {
          g = (vc == 0) ? f0 : (vc == 1) ? f1 : f2;
        }
        const bool bad = isnan(f0) || isnan(f1) || isnan(f2);  // NaN guard (:559)
        gb = bad ? 0.0 : g;
        results[p].grf[3 * ls + vc] = (float)gb;
      }
      if (tin != nullptr) {
        // fused torque map (A1RobotControl.cpp:289-319): the three variable lanes of a leg each
        // write one joint torque of that leg
        const double gx = shfl(gb, lbase), gy = shfl(gb, lbase + 2), gz = shfl(gb, lbase + 4);
        bool tnan = false;
        if (vown && lg < 4) {
          const MpcTorqueIn& t = tin[p];
          const bool contact = reinterpret_cast<const float*>(states + p)[kOffContacts + lg] != 0.0f;
          double tau[3];
          leg_torque(t.j_foot + 9 * lg, contact, gx, gy, gz, t.foot_forces_kin + 3 * lg, t.km_foot,
                     t.torques_gravity + 3 * lg, tau);
          const double tv = (vc == 0) ? tau[0] : (vc == 1) ? tau[1] : tau[2];
          tnan = isnan(tv);
          tout[p].joint_torques[3 * lg + vc] = tnan ? 0.0f : (float)tv;
        }
        // lane 6 lg + 2 vc carries component 3 lg + vc: compress the ballot to 12 bits
        const unsigned bal = __ballot_sync(0xffffffffu, tnan);
        if (lane == 0) {
          unsigned m = 0;
#pragma unroll
          for (int i = 0; i < 12; ++i) m |= ((bal >> (6 * (i / 3) + 2 * (i % 3))) & 1u) << i;
          tout[p].nan_mask = (int32_t)m;
        }
      }
    }
    if (tid == 0) {
      results[p].status = sm.flags[4] ? MPC_STATUS_INTERNAL_ERROR : status;
      results[p].iters = iter;
      results[p].rho_updates = rho_updates;
      results[p].pri_res = (float)pri_res_out;
    }
    PHASE_MARK(4);
  }
  if (kProfile && tid == 0 && phase_clk != nullptr) {
#pragma unroll
    for (int i = 0; i < (kProfile ? 6 : 1); ++i) phase_clk[blockIdx.x * 6 + i] = pc[i];
    for (int i = 0; i < 16; ++i) phase_clk[gridDim.x * 6 + blockIdx.x * 16 + i] = sm.fine[i];
  }
#undef PHASE_MARK
}

}  // namespace mpcb200
