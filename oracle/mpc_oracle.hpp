// oracle/mpc_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference hot path
//   A1RobotControl::compute_grf -> ConvexMpc -> OsqpEigen/OSQP
// used ONLY as the parity checker (tests/, __graft_entry__.smoke()) and as the
// timed CPU baseline (bench.py cpu_baseline / --impl reference).  Nothing under
// go1_qp_mpc_controller_b200/ may include, link or call this.
//
// PARITY UNPINNED: the reference stores no expected outputs for this path
// (src/a1_cpp/src/test/test_mpc.cpp only prints), and neither the reference
// (needs Eigen + OsqpEigen + ROS) nor the solver it calls can be built here.
// The solver arithmetic lives in un-vendored third-party code:
//   OSQP      -- `git clone --recursive https://github.com/oxfordcontrol/osqp`,
//                no tag pinned (docker/Dockerfile:77); workspace API => 0.6.x
//   osqp-eigen -- unpinned clone (docker/Dockerfile:91), log shows 0.6.3 (:98)
// This file restates OSQP 0.6.x's published algorithm (Stellato et al., Math.
// Prog. Comp. 2020, Alg. 1 + Alg. 2 + sec. 5.2) as the reference's call sites
// configure it (A1RobotControl.cpp:416-439, :522-555; test_mpc.cpp:131-151).
// What pins it instead: structural identities of the QP build, a KKT optimality
// certificate of the tight-tolerance solution, and an independent numpy
// restatement (tests/test_oracle.py).
//
// Everything is templated on the scalar so the SAME code can be run in double
// (the oracle) and in float (an arithmetic model of the GPU path, used by tests
// to separate "different algorithm" from "fp32 rounding").
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

constexpr int kStateDim = 13;   // MPC_STATE_DIM   A1Params.h:27
constexpr int kNumDof = 12;     // NUM_DOF         A1Params.h:34
constexpr int kConDim = 20;     // MPC_CONSTRAINT_DIM A1Params.h:28
constexpr int kNumLeg = 4;      // NUM_LEG         A1Params.h:31
constexpr double kOsqpInfty = 1e30;  // OSQP_INFTY == OsqpEigen::INFTY

// ---------------------------------------------------------------------------
// small dense helpers (row-major)
// ---------------------------------------------------------------------------
template <class T>
inline void mat3_mul(const T* a, const T* b, T* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T s = 0;
      for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
      c[3 * i + j] = s;
    }
}
template <class T>
inline void mat3_transpose(const T* a, T* t) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * j + i];
}
// Eigen's fixed 3x3 inverse() is the cofactor formula.
template <class T>
inline void mat3_inverse(const T* a, T* inv) {
  T c00 = a[4] * a[8] - a[5] * a[7];
  T c01 = a[5] * a[6] - a[3] * a[8];
  T c02 = a[3] * a[7] - a[4] * a[6];
  T det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  T id = T(1) / det;
  inv[0] = c00 * id;
  inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  inv[3] = c01 * id;
  inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  inv[6] = c02 * id;
  inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}
// Utils::skew, utils/Utils.cpp:35-41
template <class T>
inline void skew(const T* v, T* s) {
  s[0] = 0;     s[1] = -v[2]; s[2] = v[1];
  s[3] = v[2];  s[4] = 0;     s[5] = -v[0];
  s[6] = -v[1]; s[7] = v[0];  s[8] = 0;
}

// ---------------------------------------------------------------------------
// ConvexMpc restated (src/a1_cpp/src/ConvexMpc.{h,cpp}); runtime horizon.
// Public data members on purpose: the reference's callers read and write them
// (ConvexMpc.h:37, A1RobotControl.cpp:513, :527-531).
// ---------------------------------------------------------------------------
template <class T>
struct ConvexMpc {
  int H, n, s, m;
  T mu, fz_min, fz_max;
  std::vector<T> Qdiag, Rdiag;  // Q.diagonal(), R.diagonal()
  T A_mat_c[kStateDim * kStateDim];
  T B_mat_c[kStateDim * kNumDof];
  T A_mat_d[kStateDim * kStateDim];
  T B_mat_d[kStateDim * kNumDof];
  std::vector<T> B_mat_d_list;  // s x 12
  std::vector<T> A_qp;          // s x 13
  std::vector<T> B_qp;          // s x n
  std::vector<T> hessian;       // n x n (dense; the reference's sparseView drops exact zeros only)
  std::vector<T> gradient;      // n
  std::vector<T> lb, ub;        // m
  // linear_constraints as CSR triplets (ConvexMpc.cpp:46-58)
  std::vector<int> Ac_row_ptr, Ac_col;
  std::vector<T> Ac_val;

  // ConvexMpc.cpp:7-68
  ConvexMpc(const T* q_weights, const T* r_weights, int horizon)
      : H(horizon), n(kNumDof * horizon), s(kStateDim * horizon), m(kConDim * horizon) {
    mu = T(0.3);
    fz_min = 0;
    fz_max = 0;
    Qdiag.resize(s);
    Rdiag.resize(n);
    for (int i = 0; i < H; ++i)
      for (int k = 0; k < kStateDim; ++k) Qdiag[i * kStateDim + k] = 2 * q_weights[k];  // :20
    for (int i = 0; i < H; ++i)
      for (int k = 0; k < kNumDof; ++k) Rdiag[i * kNumDof + k] = 2 * r_weights[k];  // :41
    B_mat_d_list.resize(size_t(s) * kNumDof);
    A_qp.resize(size_t(s) * kStateDim);
    B_qp.resize(size_t(s) * n);
    hessian.resize(size_t(n) * n);
    gradient.resize(n);
    lb.resize(m);
    ub.resize(m);
    build_constraints();
    reset();
  }

  void build_constraints() {
    // rows 5i..5i+4 of leg-step i act on columns 3i..3i+2 (ConvexMpc.cpp:46-58)
    Ac_row_ptr.assign(1, 0);
    Ac_col.clear();
    Ac_val.clear();
    for (int i = 0; i < kNumLeg * H; ++i) {
      const int c = 3 * i;
      auto push = [&](int col, T v) { Ac_col.push_back(col); Ac_val.push_back(v); };
      push(c + 0, 1); push(c + 2, mu);  Ac_row_ptr.push_back((int)Ac_col.size());
      push(c + 0, 1); push(c + 2, -mu); Ac_row_ptr.push_back((int)Ac_col.size());
      push(c + 1, 1); push(c + 2, mu);  Ac_row_ptr.push_back((int)Ac_col.size());
      push(c + 1, 1); push(c + 2, -mu); Ac_row_ptr.push_back((int)Ac_col.size());
      push(c + 2, 1);                   Ac_row_ptr.push_back((int)Ac_col.size());
    }
  }

  // ConvexMpc.cpp:70-108
  void reset() {
    std::fill(A_mat_c, A_mat_c + kStateDim * kStateDim, T(0));
    std::fill(B_mat_c, B_mat_c + kStateDim * kNumDof, T(0));
    std::fill(A_mat_d, A_mat_d + kStateDim * kStateDim, T(0));
    std::fill(B_mat_d, B_mat_d + kStateDim * kNumDof, T(0));
    std::fill(B_mat_d_list.begin(), B_mat_d_list.end(), T(0));
    std::fill(A_qp.begin(), A_qp.end(), T(0));
    std::fill(B_qp.begin(), B_qp.end(), T(0));
    std::fill(gradient.begin(), gradient.end(), T(0));
    std::fill(lb.begin(), lb.end(), T(0));
    std::fill(ub.begin(), ub.end(), T(0));
  }

  // ConvexMpc.cpp:110-130 -- yaw only; pitch terms are commented out upstream.
  void calculate_A_mat_c(const T* root_euler) {
    T cy = std::cos(root_euler[2]);
    T sy = std::sin(root_euler[2]);
    auto A = [&](int r, int c) -> T& { return A_mat_c[r * kStateDim + c]; };
    A(0, 6) = cy;  A(0, 7) = sy; A(0, 8) = 0;
    A(1, 6) = -sy; A(1, 7) = cy; A(1, 8) = 0;
    A(2, 6) = 0;   A(2, 7) = 0;  A(2, 8) = 1;
    A(3, 9) = 1; A(4, 10) = 1; A(5, 11) = 1;
    A(11, kNumDof) = 1;
  }

  // ConvexMpc.cpp:132-143; foot_pos is 3x4 column-per-leg given here leg-major.
  void calculate_B_mat_c(T robot_mass, const T* trunk_inertia, const T* root_rot_mat,
                         const T* foot_pos_leg_major) {
    T Rt[9], tmp[9], Iw[9], Iw_inv[9];
    mat3_transpose(root_rot_mat, Rt);
    mat3_mul(root_rot_mat, trunk_inertia, tmp);
    mat3_mul(tmp, Rt, Iw);
    for (int i = 0; i < kNumLeg; ++i) {
      mat3_inverse(Iw, Iw_inv);  // the reference re-inverts per leg; same value
      T sk[9], blk[9];
      skew(foot_pos_leg_major + 3 * i, sk);
      mat3_mul(Iw_inv, sk, blk);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          B_mat_c[(6 + r) * kNumDof + 3 * i + c] = blk[3 * r + c];
          B_mat_c[(9 + r) * kNumDof + 3 * i + c] = (r == c) ? (T(1) / robot_mass) : T(0);
        }
    }
  }

  // ConvexMpc.cpp:145-156 -- forward Euler (matrix exp is commented out upstream).
  void state_space_discretization(T dt) {
    for (int r = 0; r < kStateDim; ++r)
      for (int c = 0; c < kStateDim; ++c)
        A_mat_d[r * kStateDim + c] = ((r == c) ? T(1) : T(0)) + A_mat_c[r * kStateDim + c] * dt;
    for (int i = 0; i < kStateDim * kNumDof; ++i) B_mat_d[i] = B_mat_c[i] * dt;
  }

  // SURVEY.md 8f row 4 (NOT the reference's behaviour; behind MpcConfig.exact_discretization):
  // the matrix exponential the authors left commented out (ConvexMpc.cpp:149),
  // exp([[A_c, B_c], [0, 0]] dt) = [[A_d, B_d], [0, I]], by its Taylor series (A_c is nilpotent,
  // the series terminates; twelve terms are summed without assuming that).
  void state_space_discretization_exact(T dt) {
    const int S = kStateDim, D = kNumDof, Nn = S + D;
    std::vector<T> Mx(size_t(Nn) * Nn, T(0)), term(size_t(Nn) * Nn, T(0)), acc(size_t(Nn) * Nn, T(0)), nxt(size_t(Nn) * Nn);
    for (int r = 0; r < S; ++r) {
      for (int c = 0; c < S; ++c) Mx[size_t(r) * Nn + c] = A_mat_c[r * S + c] * dt;
      for (int c = 0; c < D; ++c) Mx[size_t(r) * Nn + S + c] = B_mat_c[r * D + c] * dt;
    }
    for (int i = 0; i < Nn; ++i) term[size_t(i) * Nn + i] = acc[size_t(i) * Nn + i] = T(1);
    for (int k = 1; k <= 12; ++k) {
      for (int r = 0; r < Nn; ++r)
        for (int c = 0; c < Nn; ++c) {
          T s_ = 0;
          for (int j = 0; j < Nn; ++j) s_ += term[size_t(r) * Nn + j] * Mx[size_t(j) * Nn + c];
          nxt[size_t(r) * Nn + c] = s_ / T(k);
        }
      term = nxt;
      for (size_t i = 0; i < acc.size(); ++i) acc[i] += term[i];
    }
    for (int r = 0; r < S; ++r) {
      for (int c = 0; c < S; ++c) A_mat_d[r * S + c] = acc[size_t(r) * Nn + c];
      for (int c = 0; c < D; ++c) B_mat_d[r * D + c] = acc[size_t(r) * Nn + S + c];
    }
  }

  // The caller stores B_mat_d into B_mat_d_list block i (A1RobotControl.cpp:513).
  void store_B_mat_d(int i) {
    std::copy(B_mat_d, B_mat_d + kStateDim * kNumDof,
              B_mat_d_list.begin() + size_t(i) * kStateDim * kNumDof);
  }

  // ConvexMpc.cpp:158-245
  // contacts: 4 flags replicated over the horizon (the reference), or, with per_step, 4 H flags
  void calculate_qp_mats(const T* mpc_states, const T* mpc_states_d, const bool* contacts, bool per_step = false) {
    const int S = kStateDim, D = kNumDof;
    // :184-202  A_qp block i = A_qp block (i-1) * A_d ; B_qp block (i,j)
    for (int i = 0; i < H; ++i) {
      T* Ai = &A_qp[size_t(i) * S * S];
      if (i == 0) {
        std::copy(A_mat_d, A_mat_d + S * S, Ai);
      } else {
        const T* Ap = &A_qp[size_t(i - 1) * S * S];
        for (int r = 0; r < S; ++r)
          for (int c = 0; c < S; ++c) {
            T acc = 0;
            for (int k = 0; k < S; ++k) acc += Ap[r * S + k] * A_mat_d[k * S + c];
            Ai[r * S + c] = acc;
          }
      }
      for (int j = 0; j < i + 1; ++j) {
        const T* Bj = &B_mat_d_list[size_t(j) * S * D];
        if (i - j == 0) {
          for (int r = 0; r < S; ++r)
            for (int c = 0; c < D; ++c) B_qp[size_t(i * S + r) * n + j * D + c] = Bj[r * D + c];
        } else {
          const T* Ap = &A_qp[size_t(i - j - 1) * S * S];
          for (int r = 0; r < S; ++r)
            for (int c = 0; c < D; ++c) {
              T acc = 0;
              for (int k = 0; k < S; ++k) acc += Ap[r * S + k] * Bj[k * D + c];
              B_qp[size_t(i * S + r) * n + j * D + c] = acc;
            }
        }
      }
    }
    // :207-211  dense_hessian = B_qp' * Q * B_qp ; += R
    std::vector<T> QB(size_t(s) * n);
    for (int r = 0; r < s; ++r)
      for (int c = 0; c < n; ++c) QB[size_t(r) * n + c] = Qdiag[r] * B_qp[size_t(r) * n + c];
    std::fill(hessian.begin(), hessian.end(), T(0));
    for (int k = 0; k < s; ++k) {
      const T* bk = &B_qp[size_t(k) * n];
      const T* qk = &QB[size_t(k) * n];
      for (int r = 0; r < n; ++r) {
        const T b = bk[r];
        if (b == T(0)) continue;
        T* hr = &hessian[size_t(r) * n];
        for (int c = 0; c < n; ++c) hr[c] += b * qk[c];
      }
    }
    for (int r = 0; r < n; ++r) hessian[size_t(r) * n + r] += Rdiag[r];
    // :215-217  gradient = B_qp' * Q * (A_qp * x0 - x_ref)
    std::vector<T> tmp(s);
    for (int r = 0; r < s; ++r) {
      T acc = 0;
      for (int k = 0; k < S; ++k) acc += A_qp[size_t(r) * S + k] * mpc_states[k];
      tmp[r] = Qdiag[r] * (acc - mpc_states_d[r]);
    }
    for (int c = 0; c < n; ++c) {
      T acc = 0;
      for (int r = 0; r < s; ++r) acc += B_qp[size_t(r) * n + c] * tmp[r];
      gradient[c] = acc;
    }
    // :223-245  bounds, same contacts replicated over the horizon
    fz_min = 0;
    fz_max = 180;
    for (int h = 0; h < H; ++h)
      for (int i = 0; i < kNumLeg; ++i) {
        const T c = contacts[per_step ? 4 * h + i : i] ? T(1) : T(0);
        T* l = &lb[h * kConDim + 5 * i];
        T* u = &ub[h * kConDim + 5 * i];
        l[0] = 0;               u[0] = T(kOsqpInfty);
        l[1] = -T(kOsqpInfty);  u[1] = 0;
        l[2] = 0;               u[2] = T(kOsqpInfty);
        l[3] = -T(kOsqpInfty);  u[3] = 0;
        l[4] = fz_min * c;      u[4] = fz_max * c;
      }
  }
};

// ---------------------------------------------------------------------------
// OSQP 0.6.x restated.  Dense symmetric P (row-major n x n), CSR A (m x n).
// ---------------------------------------------------------------------------
struct Settings {
  double rho = 0.1, sigma = 1e-6, alpha = 1.6;
  double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
  int max_iter = 4000, check_termination = 25, scaling = 10;
  int adaptive_rho = 1, adaptive_rho_interval = 50;  // library default 0 = wall-clock; pinned
  double adaptive_rho_tolerance = 5.0;
};

struct Info {
  int status = -10;  // OSQP_UNSOLVED
  int iters = 0;
  int rho_updates = 0;
  double pri_res = 0, dua_res = 0, rho = 0, obj = 0;
};

// osqp constants.h
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;

template <class T>
struct Csr {
  int m = 0, n = 0;
  std::vector<int> row_ptr, col;
  std::vector<T> val;
};

template <class T>
inline T norm_inf(const std::vector<T>& v) {
  T r = 0;
  for (T x : v) r = std::max(r, std::abs(x));
  return r;
}
template <class T>
inline T limit_scaling(T v) {  // scaling.c limit_scaling
  v = v < T(kMinScaling) ? T(1) : v;
  v = v > T(kMaxScaling) ? T(kMaxScaling) : v;
  return v;
}

template <class T>
struct Osqp {
  int n, m;
  Settings st;
  std::vector<T> P, q, l, u;  // scaled in place by setup()
  Csr<T> A;
  Csr<T> A0;                  // unscaled constraint matrix (for osqp_update_P's unscale/scale)
  std::vector<T> q_unscaled;  // the linear cost as last passed by the caller
  std::vector<T> D, E, Dinv, Einv;
  T c = 1, cinv = 1;
  std::vector<T> rho_vec, rho_inv_vec;
  std::vector<int> constr_type;
  std::vector<T> L;  // Cholesky factor of K = P + sigma I + A' diag(rho) A, row-major lower
  std::vector<T> x, z, y, x_prev, z_prev, xt, zt, delta_y, delta_x, Ax, Px, Aty;
  Info info;
  int n_factor = 0;

  void Amul(const std::vector<T>& v, std::vector<T>& out) const {
    for (int i = 0; i < m; ++i) {
      T sacc = 0;
      for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) sacc += A.val[k] * v[A.col[k]];
      out[i] = sacc;
    }
  }
  void Atmul(const std::vector<T>& v, std::vector<T>& out) const {
    std::fill(out.begin(), out.end(), T(0));
    for (int i = 0; i < m; ++i)
      for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) out[A.col[k]] += A.val[k] * v[i];
  }
  void Pmul(const std::vector<T>& v, std::vector<T>& out) const {
    for (int i = 0; i < n; ++i) {
      T sacc = 0;
      const T* pr = &P[size_t(i) * n];
      for (int j = 0; j < n; ++j) sacc += pr[j] * v[j];
      out[i] = sacc;
    }
  }

  // scaling.c scale_data: modified Ruiz on [[P,A'],[A,0]] with the cost
  // normalisation INSIDE the loop (paper Alg. 2).
  void scale_data() {
    D.assign(n, T(1));
    E.assign(m, T(1));
    c = 1;
    std::vector<T> Dt(n), Et(m);
    for (int it = 0; it < st.scaling; ++it) {
      // column inf-norms of the KKT matrix
      for (int j = 0; j < n; ++j) Dt[j] = 0;
      for (int i = 0; i < n; ++i) {
        const T* pr = &P[size_t(i) * n];
        for (int j = 0; j < n; ++j) Dt[j] = std::max(Dt[j], std::abs(pr[j]));
      }
      for (int i = 0; i < m; ++i) {
        T rmax = 0;
        for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
          T a = std::abs(A.val[k]);
          rmax = std::max(rmax, a);
          Dt[A.col[k]] = std::max(Dt[A.col[k]], a);
        }
        Et[i] = rmax;
      }
      for (int j = 0; j < n; ++j) Dt[j] = T(1) / std::sqrt(limit_scaling(Dt[j]));
      for (int i = 0; i < m; ++i) Et[i] = T(1) / std::sqrt(limit_scaling(Et[i]));
      // P <- D P D, A <- E A D, q <- D q
      for (int i = 0; i < n; ++i) {
        T* pr = &P[size_t(i) * n];
        for (int j = 0; j < n; ++j) pr[j] = Dt[i] * pr[j] * Dt[j];
      }
      for (int i = 0; i < m; ++i)
        for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) A.val[k] = Et[i] * A.val[k] * Dt[A.col[k]];
      for (int j = 0; j < n; ++j) q[j] *= Dt[j];
      for (int j = 0; j < n; ++j) D[j] *= Dt[j];
      for (int i = 0; i < m; ++i) E[i] *= Et[i];
      // cost normalisation
      T mean = 0;
      for (int j = 0; j < n; ++j) {
        T cmax = 0;
        for (int i = 0; i < n; ++i) cmax = std::max(cmax, std::abs(P[size_t(i) * n + j]));
        mean += cmax;
      }
      mean /= T(n);
      T nq = limit_scaling(norm_inf(q));
      T ct = T(1) / limit_scaling(std::max(mean, nq));
      for (auto& v : P) v *= ct;
      for (auto& v : q) v *= ct;
      c *= ct;
    }
    cinv = T(1) / c;
    Dinv.resize(n);
    Einv.resize(m);
    for (int j = 0; j < n; ++j) Dinv[j] = T(1) / D[j];
    for (int i = 0; i < m; ++i) Einv[i] = T(1) / E[i];
    for (int i = 0; i < m; ++i) {
      l[i] *= E[i];
      u[i] *= E[i];
    }
  }

  // auxil.c set_rho_vec / update_rho_vec
  void set_rho_vec(bool first) {
    T rho = T(st.rho);
    rho = std::min(std::max(rho, T(kRhoMin)), T(kRhoMax));
    st.rho = double(rho);
    if (first) constr_type.assign(m, 0);
    rho_vec.resize(m);
    rho_inv_vec.resize(m);
    for (int i = 0; i < m; ++i) {
      if (first) {
        if (l[i] < -T(kOsqpInfty * kMinScaling) && u[i] > T(kOsqpInfty * kMinScaling))
          constr_type[i] = -1;
        else if (u[i] - l[i] < T(kRhoTol))
          constr_type[i] = 1;
        else
          constr_type[i] = 0;
      }
      if (constr_type[i] == -1) rho_vec[i] = T(kRhoMin);
      else if (constr_type[i] == 1) rho_vec[i] = T(kRhoEqOverIneq) * rho;
      else rho_vec[i] = rho;
      rho_inv_vec[i] = T(1) / rho_vec[i];
    }
  }

  // OSQP factors the quasi-definite KKT [[P+sigma I, A'],[A, -1/rho]] with
  // QDLDL; eliminating nu gives K x~ = sigma x - q + A'(rho z - y), z~ = A x~
  // with K = P + sigma I + A' diag(rho) A (identical iterates in exact
  // arithmetic).  Dense Cholesky here.
  bool factor() {
    ++n_factor;
    std::vector<T> K(P);
    for (int j = 0; j < n; ++j) K[size_t(j) * n + j] += T(st.sigma);
    for (int i = 0; i < m; ++i) {
      const T r = rho_vec[i];
      for (int a = A.row_ptr[i]; a < A.row_ptr[i + 1]; ++a)
        for (int b = A.row_ptr[i]; b < A.row_ptr[i + 1]; ++b)
          K[size_t(A.col[a]) * n + A.col[b]] += r * A.val[a] * A.val[b];
    }
    L.assign(size_t(n) * n, T(0));
    for (int j = 0; j < n; ++j) {
      T d = K[size_t(j) * n + j];
      for (int k = 0; k < j; ++k) d -= L[size_t(j) * n + k] * L[size_t(j) * n + k];
      if (!(d > T(0))) return false;
      d = std::sqrt(d);
      L[size_t(j) * n + j] = d;
      const T dinv = T(1) / d;
      for (int i = j + 1; i < n; ++i) {
        T sacc = K[size_t(i) * n + j];
        const T* li = &L[size_t(i) * n];
        const T* lj = &L[size_t(j) * n];
        for (int k = 0; k < j; ++k) sacc -= li[k] * lj[k];
        L[size_t(i) * n + j] = sacc * dinv;
      }
    }
    return true;
  }
  void chol_solve(std::vector<T>& b) const {
    for (int i = 0; i < n; ++i) {
      T sacc = b[i];
      const T* li = &L[size_t(i) * n];
      for (int k = 0; k < i; ++k) sacc -= li[k] * b[k];
      b[i] = sacc / li[i];
    }
    for (int i = n - 1; i >= 0; --i) {
      T sacc = b[i];
      for (int k = i + 1; k < n; ++k) sacc -= L[size_t(k) * n + i] * b[k];
      b[i] = sacc / L[size_t(i) * n + i];
    }
  }

  // osqp_setup: copy data, scale, rho vector, factor, cold start.
  bool setup(int n_, int m_, const T* P_, const T* q_, const Csr<T>& A_, const T* l_, const T* u_,
             const Settings& s_) {
    n = n_;
    m = m_;
    st = s_;
    P.assign(P_, P_ + size_t(n) * n);
    q.assign(q_, q_ + n);
    l.assign(l_, l_ + m);
    u.assign(u_, u_ + m);
    A = A_;
    A0 = A_;
    q_unscaled = q;
    if (st.scaling > 0) {
      scale_data();
    } else {
      D.assign(n, 1); Dinv.assign(n, 1); E.assign(m, 1); Einv.assign(m, 1);
      c = cinv = 1;
    }
    set_rho_vec(true);
    x.assign(n, 0); z.assign(m, 0); y.assign(m, 0);
    x_prev.assign(n, 0); z_prev.assign(m, 0);
    xt.assign(n, 0); zt.assign(m, 0);
    delta_y.assign(m, 0); delta_x.assign(n, 0);
    Ax.assign(m, 0); Px.assign(n, 0); Aty.assign(n, 0);
    info = Info();
    n_factor = 0;
    return factor();
  }

  // The warm path of the live controller (A1RobotControl.cpp:532-538): updateHessianMatrix,
  // updateGradient, updateLowerBound, updateUpperBound on a solver that stays alive.
  //   osqp_update_P        unscale_data; new P; scale_data (from D = E = c = 1, with the OLD q
  //                        still in place: the cost scaling c sees the previous gradient);
  //                        refactor; reset_info
  //   osqp_update_lin_cost q <- c D q_new
  //   osqp_update_bounds   l, u <- E l_new, E u_new; update_rho_vec (constraint types may change)
  // The iterates x, z, y (in the OLD scaled coordinates) and settings->rho are kept: warm start.
  bool update(const T* P_new, const T* q_new, const T* l_new, const T* u_new) {
    P.assign(P_new, P_new + size_t(n) * n);
    q = q_unscaled;  // old gradient, unscaled
    A = A0;
    if (st.scaling > 0) {
      // l, u are rescaled by scale_data and then overwritten below: give it neutral values
      std::fill(l.begin(), l.end(), T(0));
      std::fill(u.begin(), u.end(), T(0));
      scale_data();
    }
    q_unscaled.assign(q_new, q_new + n);
    for (int j = 0; j < n; ++j) q[j] = c * D[j] * q_unscaled[j];
    for (int i = 0; i < m; ++i) {
      l[i] = E[i] * l_new[i];
      u[i] = E[i] * u_new[i];
    }
    set_rho_vec(true);  // recomputes the constraint types from the new scaled bounds, keeps st.rho
    info = Info();
    return factor();
  }

  // auxil.c compute_pri_res / compute_dua_res / tolerances (scaled_termination = 0)
  struct Residuals {
    T pri, dua, pri_scaled, dua_scaled, eps_pri, eps_dua;
    T pri_norm, dua_norm;  // max(|z|,|Ax|) and c^-1 max(|q|,|A'y|,|Px|), unscaled
  };
  Residuals residuals(std::vector<T>& wm, std::vector<T>& wn) {
    Residuals r;
    Amul(x, Ax);
    T ps = 0, pu = 0;
    for (int i = 0; i < m; ++i) {
      wm[i] = Ax[i] - z[i];
      ps = std::max(ps, std::abs(wm[i]));
      pu = std::max(pu, std::abs(Einv[i] * wm[i]));
    }
    r.pri_scaled = ps;
    r.pri = pu;
    Pmul(x, Px);
    Atmul(y, Aty);
    T ds = 0, du = 0;
    for (int j = 0; j < n; ++j) {
      wn[j] = q[j] + Px[j] + Aty[j];
      ds = std::max(ds, std::abs(wn[j]));
      du = std::max(du, std::abs(Dinv[j] * wn[j]));
    }
    r.dua_scaled = ds;
    r.dua = cinv * du;
    T nz = 0, nAx = 0;
    for (int i = 0; i < m; ++i) {
      nz = std::max(nz, std::abs(Einv[i] * z[i]));
      nAx = std::max(nAx, std::abs(Einv[i] * Ax[i]));
    }
    r.pri_norm = std::max(nz, nAx);
    r.eps_pri = T(st.eps_abs) + T(st.eps_rel) * r.pri_norm;
    T nq = 0, nAty = 0, nPx = 0;
    for (int j = 0; j < n; ++j) {
      nq = std::max(nq, std::abs(Dinv[j] * q[j]));
      nAty = std::max(nAty, std::abs(Dinv[j] * Aty[j]));
      nPx = std::max(nPx, std::abs(Dinv[j] * Px[j]));
    }
    r.dua_norm = cinv * std::max(std::max(nq, nAty), nPx);
    r.eps_dua = T(st.eps_abs) + T(st.eps_rel) * r.dua_norm;
    return r;
  }

  // auxil.c is_primal_infeasible / is_dual_infeasible (scaling on, unscaled test)
  bool primal_infeasible(T eps) {
    // project delta_y onto the polar of the recession cone of [l,u]
    const T big = T(kOsqpInfty * kMinScaling);
    for (int i = 0; i < m; ++i) {
      if (u[i] > big) {
        if (l[i] < -big) delta_y[i] = 0;
        else delta_y[i] = std::min(delta_y[i], T(0));
      } else if (l[i] < -big) {
        delta_y[i] = std::max(delta_y[i], T(0));
      }
    }
    T nd = 0;
    for (int i = 0; i < m; ++i) nd = std::max(nd, std::abs(E[i] * delta_y[i]));
    if (!(nd > eps)) return false;
    T lhs = 0;
    for (int i = 0; i < m; ++i)
      lhs += u[i] * std::max(delta_y[i], T(0)) + l[i] * std::min(delta_y[i], T(0));
    if (!(lhs < -eps * nd)) return false;
    std::vector<T> w(n);
    Atmul(delta_y, w);
    T na = 0;
    for (int j = 0; j < n; ++j) na = std::max(na, std::abs(Dinv[j] * w[j]));
    return na < eps * nd;
  }
  bool dual_infeasible(T eps) {
    T nd = 0;
    for (int j = 0; j < n; ++j) nd = std::max(nd, std::abs(D[j] * delta_x[j]));
    if (!(nd > eps)) return false;
    T qdx = 0;
    for (int j = 0; j < n; ++j) qdx += q[j] * delta_x[j];
    if (!(qdx * cinv < -eps * nd)) return false;
    std::vector<T> w(n);
    Pmul(delta_x, w);
    T np = 0;
    for (int j = 0; j < n; ++j) np = std::max(np, std::abs(Dinv[j] * w[j]));
    if (!(cinv * np < eps * nd)) return false;
    std::vector<T> adx(m);
    Amul(delta_x, adx);
    for (int i = 0; i < m; ++i) {
      const T v = Einv[i] * adx[i];
      const bool u_inf = u[i] > T(kOsqpInfty * kMinScaling);
      const bool l_inf = l[i] < -T(kOsqpInfty * kMinScaling);
      if ((!u_inf && v > eps * nd) || (!l_inf && v < -eps * nd)) return false;
    }
    return true;
  }

  // osqp_solve main loop (osqp.c), polish off, scaled_termination off.
  void solve() {
    std::vector<T> rhs(n), w(m), wm(m), wn(n);
    const T sigma = T(st.sigma), alpha = T(st.alpha);
    int iter = 0;
    bool checked_last = false;
    for (iter = 1; iter <= st.max_iter; ++iter) {
      x_prev.swap(x);
      z_prev.swap(z);
      // update_xz_tilde: K x~ = sigma x_prev - q + A'(rho z_prev - y); z~ = A x~
      for (int i = 0; i < m; ++i) w[i] = rho_vec[i] * z_prev[i] - y[i];
      Atmul(w, rhs);
      for (int j = 0; j < n; ++j) rhs[j] += sigma * x_prev[j] - q[j];
      chol_solve(rhs);
      xt = rhs;
      Amul(xt, zt);
      // update_x
      for (int j = 0; j < n; ++j) {
        x[j] = alpha * xt[j] + (T(1) - alpha) * x_prev[j];
        delta_x[j] = x[j] - x_prev[j];
      }
      // update_z, update_y
      for (int i = 0; i < m; ++i) {
        const T zr = alpha * zt[i] + (T(1) - alpha) * z_prev[i];
        T zn = zr + rho_inv_vec[i] * y[i];
        zn = std::min(std::max(zn, l[i]), u[i]);
        z[i] = zn;
        delta_y[i] = rho_vec[i] * (zr - zn);
        y[i] += delta_y[i];
      }
      checked_last = false;
      const bool can_check = st.check_termination && (iter % st.check_termination == 0);
      const bool can_adapt = st.adaptive_rho && st.adaptive_rho_interval &&
                             (iter % st.adaptive_rho_interval == 0);
      Residuals r{};
      if (can_check || can_adapt) {
        r = residuals(wm, wn);
        info.pri_res = double(r.pri);
        info.dua_res = double(r.dua);
      }
      if (can_check) {
        checked_last = true;
        if (check_termination(r, false)) break;
      }
      if (can_adapt) {
        // compute_rho_estimate (scaled quantities)
        T pn = std::max(norm_inf(z), norm_inf(Ax));
        T dn = std::max(std::max(norm_inf(q), norm_inf(Aty)), norm_inf(Px));
        T pr = r.pri_scaled / (pn + T(1e-10));
        T dr = r.dua_scaled / (dn + T(1e-10));
        T rho_new = T(st.rho) * std::sqrt(pr / (dr + T(1e-10)));
        rho_new = std::min(std::max(rho_new, T(kRhoMin)), T(kRhoMax));
        if (rho_new > T(st.rho) * T(st.adaptive_rho_tolerance) ||
            rho_new < T(st.rho) / T(st.adaptive_rho_tolerance)) {
          st.rho = double(rho_new);
          set_rho_vec(false);
          factor();
          ++info.rho_updates;
        }
      }
    }
    if (iter > st.max_iter) iter = st.max_iter;
    info.iters = iter;
    if (!checked_last) {
      Residuals r = residuals(wm, wn);
      info.pri_res = double(r.pri);
      info.dua_res = double(r.dua);
      check_termination(r, false);
    }
    if (info.status == -10) {
      // osqp.c: "if max iterations reached, try the approximate check"
      Residuals r = residuals(wm, wn);
      if (!check_termination(r, true)) info.status = -2;
    }
    info.rho = st.rho;
  }

  // auxil.c check_termination; approximate = tolerances x 10 -> *_INACCURATE
  bool check_termination(const Residuals& r, bool approximate) {
    const T k = approximate ? T(10) : T(1);
    const T eps_pri = k * (T(st.eps_abs) + T(st.eps_rel) * r.pri_norm);
    const T eps_dua = k * (T(st.eps_abs) + T(st.eps_rel) * r.dua_norm);
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0 || r.pri < eps_pri) prim_ok = true;
    else prim_inf = primal_infeasible(k * T(st.eps_prim_inf));
    if (r.dua < eps_dua) dual_ok = true;
    else dual_inf = dual_infeasible(k * T(st.eps_dual_inf));
    if (prim_ok && dual_ok) { info.status = approximate ? 2 : 1; return true; }
    if (prim_inf) { info.status = approximate ? 3 : -3; return true; }
    if (dual_inf) { info.status = approximate ? 4 : -4; return true; }
    return false;
  }

  // store_solution + unscale_solution
  void solution(T* x_out, T* y_out) const {
    const bool ok = info.status == 1 || info.status == -2 || info.status == 2;
    for (int j = 0; j < n; ++j) x_out[j] = ok ? D[j] * x[j] : std::nan("");
    if (y_out)
      for (int i = 0; i < m; ++i) y_out[i] = ok ? cinv * E[i] * y[i] : std::nan("");
  }
};

// ---------------------------------------------------------------------------
// compute_grf restated (A1RobotControl.cpp:321-564), one robot.
// ---------------------------------------------------------------------------
struct MpcParams {
  int H = 10;
  double dt = 0.0025, mu = 0.3, fz_min = 0, fz_max = 180, mass = 12.0;
  double inertia[9] = {0.0168352186, 0, 0, 0, 0.0656071082, 0, 0, 0, 0.0742720659};
  double q_weights[13] = {20, 10, 1, 0, 0, 420, 0.05, 0.05, 0.05, 30, 30, 10, 0};
  double r_weights[12] = {1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7};
  Settings osqp;
  // SURVEY.md 8f row 4 extensions (change results; all false = the reference)
  bool exact_discretization = false, foot_drift = false, gait_aware = false;
};

// Field-for-field the MpcStateIn record of include/mpc_b200.h, already widened.
template <class T>
struct RobotState {
  T euler[3], pos[3], ang_vel[3], lin_vel[3], euler_d[3], pos_d_z, lin_vel_d[3], ang_vel_d[3];
  T rot_mat[9], foot_pos_abs[12];
  bool contacts[4];
};

template <class T>
struct MpcProblem {
  std::vector<T> x0, x_ref;
  ConvexMpc<T> mpc;
  MpcProblem(const MpcParams& p)
      : x0(kStateDim), x_ref(size_t(kStateDim) * p.H), mpc(to_T(p.q_weights, 13).data(),
                                                           to_T(p.r_weights, 12).data(), p.H) {}
  static std::vector<T> to_T(const double* v, int k) {
    std::vector<T> o(k);
    for (int i = 0; i < k; ++i) o[i] = T(v[i]);
    return o;
  }
};

// A1RobotControl.cpp:446-518: pack x0 / reference, drive ConvexMpc.
template <class T>
void mpc_build(const MpcParams& p, const RobotState<T>& st, MpcProblem<T>& pb, const MpcGaitIn* gait = nullptr) {
  ConvexMpc<T>& mpc = pb.mpc;
  mpc.reset();                                   // :448
  mpc.mu = T(p.mu);
  mpc.build_constraints();
  T* x0 = pb.x0.data();                          // :452-456
  for (int i = 0; i < 3; ++i) {
    x0[i] = st.euler[i];
    x0[3 + i] = st.pos[i];
    x0[6 + i] = st.ang_vel[i];
    x0[9 + i] = st.lin_vel[i];
  }
  x0[12] = T(-9.8);
  const T dt = T(p.dt);                          // :462
  T vdw[3];                                      // :470 root_lin_vel_d_world = R * v_d
  for (int r = 0; r < 3; ++r)
    vdw[r] = st.rot_mat[3 * r] * st.lin_vel_d[0] + st.rot_mat[3 * r + 1] * st.lin_vel_d[1] +
             st.rot_mat[3 * r + 2] * st.lin_vel_d[2];
  for (int i = 0; i < p.H; ++i) {                // :472-488
    T* d = &pb.x_ref[size_t(i) * kStateDim];
    d[0] = st.euler_d[0];
    d[1] = st.euler_d[1];
    d[2] = st.euler[2] + st.ang_vel_d[2] * dt * T(i + 1);
    d[3] = st.pos[0] + vdw[0] * dt * T(i + 1);
    d[4] = st.pos[1] + vdw[1] * dt * T(i + 1);
    d[5] = st.pos_d_z;
    d[6] = st.ang_vel_d[0];
    d[7] = st.ang_vel_d[1];
    d[8] = st.ang_vel_d[2];
    d[9] = vdw[0];
    d[10] = vdw[1];
    d[11] = 0;
    d[12] = T(-9.8);
  }
  mpc.calculate_A_mat_c(st.euler);               // :492
  T inertia[9];
  for (int i = 0; i < 9; ++i) inertia[i] = T(p.inertia[i]);
  for (int i = 0; i < p.H; ++i) {                // :498-514 (same B_d every step)
    T foot[12];
    for (int k = 0; k < 12; ++k) foot[k] = st.foot_pos_abs[k];
    if (p.foot_drift)                            // the update commented out at :504-507, in the world frame
      for (int leg = 0; leg < kNumLeg; ++leg)
        for (int r = 0; r < 3; ++r) foot[3 * leg + r] -= T(i) * dt * vdw[r];
    mpc.calculate_B_mat_c(T(p.mass), inertia, st.rot_mat, foot);
    if (p.exact_discretization) mpc.state_space_discretization_exact(dt);
    else mpc.state_space_discretization(dt);
    mpc.store_B_mat_d(i);
  }
  if (p.gait_aware && gait) {
    // planned contacts of step i >= 1 from the gait counters (A1RobotControl.cpp:156-164)
    std::vector<char> cs(size_t(4) * p.H);
    bool* flags = reinterpret_cast<bool*>(cs.data());
    for (int i = 0; i < p.H; ++i)
      for (int leg = 0; leg < kNumLeg; ++leg) {
        bool c = st.contacts[leg];
        if (i > 0) {
          const double cnt = std::fmod(double(gait->gait_counter[leg]) +
                                           double(i) * double(gait->ticks_per_step) * double(gait->gait_counter_speed[leg]),
                                       double(gait->counter_per_gait));
          c = cnt <= double(gait->counter_per_swing);
        }
        flags[4 * i + leg] = c;
      }
    mpc.calculate_qp_mats(x0, pb.x_ref.data(), flags, true);
  } else {
    mpc.calculate_qp_mats(x0, pb.x_ref.data(), st.contacts);  // :518
  }
}

template <class T>
struct GrfResult {
  T grf[12];
  int status, iters, rho_updates;
  double pri_res, dua_res;
};

// A1RobotControl.cpp:522-561: cold-start solve, first step rotated by R'.
template <class T>
void mpc_solve(const MpcParams& p, const RobotState<T>& st, MpcProblem<T>& pb, GrfResult<T>& out,
               T* full_solution = nullptr) {
  ConvexMpc<T>& mpc = pb.mpc;
  Csr<T> A;
  A.m = mpc.m;
  A.n = mpc.n;
  A.row_ptr = mpc.Ac_row_ptr;
  A.col = mpc.Ac_col;
  A.val = mpc.Ac_val;
  Osqp<T> solver;
  solver.setup(mpc.n, mpc.m, mpc.hessian.data(), mpc.gradient.data(), A, mpc.lb.data(),
               mpc.ub.data(), p.osqp);
  solver.solve();
  std::vector<T> sol(mpc.n);
  solver.solution(sol.data(), nullptr);
  if (full_solution) std::copy(sol.begin(), sol.end(), full_solution);
  for (int i = 0; i < kNumLeg; ++i) {
    const T* f = &sol[3 * i];
    const T nrm = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int r = 0; r < 3; ++r) {
      // R' * f ; on NaN the reference leaves the entry untouched (:559), the
      // engine contract writes zero.
      out.grf[3 * i + r] = std::isnan(nrm) ? T(0)
                                           : st.rot_mat[r] * f[0] + st.rot_mat[3 + r] * f[1] +
                                                 st.rot_mat[6 + r] * f[2];
    }
  }
  out.status = solver.info.status;
  out.iters = solver.info.iters;
  out.rho_updates = solver.info.rho_updates;
  out.pri_res = solver.info.pri_res;
  out.dua_res = solver.info.dua_res;
}

// One robot's persistent MPC solver over control ticks (A1RobotControl.h:67 `solver` member,
// A1RobotControl.cpp:522-540): initSolver on the first tick, update* + warm solve afterwards.
template <class T>
struct MpcStream {
  Osqp<T> solver;
  bool initialised = false;
  void tick(const MpcParams& p, const RobotState<T>& st, MpcProblem<T>& pb, GrfResult<T>& out) {
    ConvexMpc<T>& mpc = pb.mpc;
    if (!initialised) {
      Csr<T> A;
      A.m = mpc.m; A.n = mpc.n;
      A.row_ptr = mpc.Ac_row_ptr; A.col = mpc.Ac_col; A.val = mpc.Ac_val;
      solver.setup(mpc.n, mpc.m, mpc.hessian.data(), mpc.gradient.data(), A, mpc.lb.data(), mpc.ub.data(), p.osqp);
      initialised = true;
    } else {
      solver.update(mpc.hessian.data(), mpc.gradient.data(), mpc.lb.data(), mpc.ub.data());
    }
    solver.solve();
    std::vector<T> sol(mpc.n);
    solver.solution(sol.data(), nullptr);
    for (int i = 0; i < kNumLeg; ++i) {
      const T* f = &sol[3 * i];
      const bool bad = std::isnan(f[0]) || std::isnan(f[1]) || std::isnan(f[2]);
      for (int r = 0; r < 3; ++r)
        out.grf[3 * i + r] = bad ? T(0) : st.rot_mat[r] * f[0] + st.rot_mat[3 + r] * f[1] + st.rot_mat[6 + r] * f[2];
    }
    out.status = solver.info.status;
    out.iters = solver.info.iters;
    out.rho_updates = solver.info.rho_updates;
    out.pri_res = solver.info.pri_res;
    out.dua_res = solver.info.dua_res;
  }
};

// ---------------------------------------------------------------------------
// compute_joint_torques (A1RobotControl.cpp:289-319) for one robot.
//   j_foot      : the four 3x3 diagonal blocks, leg-major, row-major
//   grf, f_kin  : 3 per leg, robot frame
// Returns the NaN mask (bit i: component i is NaN, the caller keeps its previous torque, :314-317).
// The swing-leg solve is Eigen's jac.lu().solve(): LU with partial (row) pivoting.
template <class T>
inline int torque_map(const T* j_foot, const bool* contacts, const T* grf, const T* f_kin, const T* km,
                      const T* torques_gravity, T* joint_torques) {
  int mask = 0;
  for (int leg = 0; leg < kNumLeg; ++leg) {
    T A[3][3], tau[3];
    for (int i = 0; i < 9; ++i) A[i / 3][i % 3] = j_foot[9 * leg + i];
    if (contacts[leg]) {
      // jac^T * -grf (:303)
      for (int k = 0; k < 3; ++k) {
        T acc = 0;
        for (int r = 0; r < 3; ++r) acc += A[r][k] * -grf[3 * leg + r];
        tau[k] = acc;
      }
    } else {
      // jac * tau = km .* f_kin (:306-307)
      T rhs[3];
      int perm[3] = {0, 1, 2};
      for (int i = 0; i < 3; ++i) rhs[i] = km[i] * f_kin[3 * leg + i];
      for (int col = 0; col < 3; ++col) {
        int best = col;
        for (int r = col + 1; r < 3; ++r)
          if (std::abs(A[perm[r]][col]) > std::abs(A[perm[best]][col])) best = r;
        std::swap(perm[col], perm[best]);
        for (int r = col + 1; r < 3; ++r) {
          const T f = A[perm[r]][col] / A[perm[col]][col];
          A[perm[r]][col] = f;
          for (int c = col + 1; c < 3; ++c) A[perm[r]][c] -= f * A[perm[col]][c];
        }
      }
      T yv[3];
      for (int i = 0; i < 3; ++i) {
        T acc = rhs[perm[i]];
        for (int c = 0; c < i; ++c) acc -= A[perm[i]][c] * yv[c];
        yv[i] = acc;
      }
      for (int i = 2; i >= 0; --i) {
        T acc = yv[i];
        for (int c = i + 1; c < 3; ++c) acc -= A[perm[i]][c] * tau[c];
        tau[i] = acc / A[perm[i]][i];
      }
    }
    for (int k = 0; k < 3; ++k) {
      const T v = tau[k] + torques_gravity[3 * leg + k];  // gravity compensation (:311)
      joint_torques[3 * leg + k] = v;
      if (std::isnan(v)) mask |= 1 << (3 * leg + k);
    }
  }
  return mask;
}

// ---------------------------------------------------------------------------
// stance-balance QP (A1RobotControl.cpp:11-48 constants, :321-332, :377-444)
// ---------------------------------------------------------------------------
struct BalanceParams {
  double Q[6] = {1.0, 1.0, 1.0, 400.0, 400.0, 100.0};
  double R = 1e-3, mu = 0.7, F_min = 0, F_max = 180, mass = 12.0;
  double kp_linear[3] = {100, 100, 300}, kd_linear[3] = {70, 70, 120};   // gazebo_a1_qp.yaml:54-60
  double kp_angular[3] = {150, 150, 1}, kd_angular[3] = {4.5, 4.5, 30};  // :62-68
  Settings osqp;
};

template <class T>
struct BalanceState {
  T euler[3], pos[3], ang_vel[3], lin_vel[3], euler_d[3], pos_d[3], lin_vel_d[3], ang_vel_d[3];
  T rot_mat[9], rot_mat_z[9], foot_pos_abs[12];
  bool contacts[4];
};

template <class T>
struct BalanceProblem {
  T P[144], q[12], l[20], u[20];
  Csr<T> A;
};

template <class T>
void balance_build(const BalanceParams& p, const BalanceState<T>& st, BalanceProblem<T>& pb) {
  T euler_error[3];
  for (int i = 0; i < 3; ++i) euler_error[i] = st.euler_d[i] - st.euler[i];
  if (euler_error[2] > T(3.1415926 * 1.5))                                 // :325-332
    euler_error[2] = st.euler_d[2] - T(3.1415926 * 2) - st.euler[2];
  else if (euler_error[2] < T(-3.1415926 * 1.5))
    euler_error[2] = st.euler_d[2] + T(3.1415926 * 2) - st.euler[2];
  const T* R = st.rot_mat;
  T root_acc[6];
  T Rtv[3], Rtw[3];
  for (int r = 0; r < 3; ++r) {
    Rtv[r] = R[r] * st.lin_vel[0] + R[3 + r] * st.lin_vel[1] + R[6 + r] * st.lin_vel[2];
    Rtw[r] = R[r] * st.ang_vel[0] + R[3 + r] * st.ang_vel[1] + R[6 + r] * st.ang_vel[2];
  }
  T tmp[3];
  for (int r = 0; r < 3; ++r) tmp[r] = T(p.kd_linear[r]) * (st.lin_vel_d[r] - Rtv[r]);
  for (int r = 0; r < 3; ++r) {
    root_acc[r] = T(p.kp_linear[r]) * (st.pos_d[r] - st.pos[r]) +                  // :381
                  (R[3 * r] * tmp[0] + R[3 * r + 1] * tmp[1] + R[3 * r + 2] * tmp[2]);  // :383-384
    root_acc[3 + r] = T(p.kp_angular[r]) * euler_error[r] +                        // :386
                      T(p.kd_angular[r]) * (st.ang_vel_d[r] - Rtw[r]);             // :388-389
  }
  root_acc[2] += T(p.mass) * T(9.8);                                               // :391
  T M[6 * 12];                                                                     // :394-399
  T Rzt[9];
  mat3_transpose(st.rot_mat_z, Rzt);
  for (int i = 0; i < kNumLeg; ++i) {
    T sk[9], blk[9];
    skew(st.foot_pos_abs + 3 * i, sk);
    mat3_mul(Rzt, sk, blk);
    for (int r = 0; r < 3; ++r)
      for (int cc = 0; cc < 3; ++cc) {
        M[r * 12 + 3 * i + cc] = (r == cc) ? T(1) : T(0);
        M[(3 + r) * 12 + 3 * i + cc] = blk[3 * r + cc];
      }
  }
  for (int a = 0; a < 12; ++a) {                                                   // :400-406
    for (int b = 0; b < 12; ++b) {
      T sacc = (a == b) ? T(p.R) : T(0);
      for (int k = 0; k < 6; ++k) sacc += M[k * 12 + a] * T(p.Q[k]) * M[k * 12 + b];
      pb.P[a * 12 + b] = sacc;
    }
    T g = 0;
    for (int k = 0; k < 6; ++k) g += M[k * 12 + a] * T(p.Q[k]) * root_acc[k];
    pb.q[a] = -g;
  }
  // constraint matrix (:28-48) and bounds (:409-413)
  Csr<T>& A = pb.A;
  A.m = 20;
  A.n = 12;
  A.row_ptr.assign(1, 0);
  A.col.clear();
  A.val.clear();
  auto push = [&](int col, T v) { A.col.push_back(col); A.val.push_back(v); };
  for (int i = 0; i < kNumLeg; ++i) {
    push(2 + 3 * i, 1);
    A.row_ptr.push_back((int)A.col.size());
    const T cflag = st.contacts[i] ? T(1) : T(0);
    pb.l[i] = cflag * T(p.F_min);
    pb.u[i] = cflag * T(p.F_max);
  }
  const T mu = T(p.mu);
  for (int i = 0; i < kNumLeg; ++i) {
    push(3 * i, 1);      push(2 + 3 * i, -mu); A.row_ptr.push_back((int)A.col.size());
    push(3 * i, -1);     push(2 + 3 * i, -mu); A.row_ptr.push_back((int)A.col.size());
    push(1 + 3 * i, 1);  push(2 + 3 * i, -mu); A.row_ptr.push_back((int)A.col.size());
    push(1 + 3 * i, -1); push(2 + 3 * i, -mu); A.row_ptr.push_back((int)A.col.size());
    for (int k = 0; k < 4; ++k) {
      pb.l[4 + 4 * i + k] = -T(kOsqpInfty);
      pb.u[4 + 4 * i + k] = 0;
    }
  }
}

template <class T>
void balance_solve(const BalanceParams& p, const BalanceState<T>& st, BalanceProblem<T>& pb,
                   GrfResult<T>& out) {
  Osqp<T> solver;
  solver.setup(12, 20, pb.P, pb.q, pb.A, pb.l, pb.u, p.osqp);
  solver.solve();
  T sol[12];
  solver.solution(sol, nullptr);
  const T* R = st.rot_mat;
  for (int i = 0; i < kNumLeg; ++i) {
    const T* f = &sol[3 * i];
    const bool bad = std::isnan(f[0]) || std::isnan(f[1]) || std::isnan(f[2]);
    for (int r = 0; r < 3; ++r)  // :439-444  R' * f
      out.grf[3 * i + r] = bad ? T(0) : R[r] * f[0] + R[3 + r] * f[1] + R[6 + r] * f[2];
  }
  out.status = solver.info.status;
  out.iters = solver.info.iters;
  out.rho_updates = solver.info.rho_updates;
  out.pri_res = solver.info.pri_res;
  out.dua_res = solver.info.dua_res;
}

}  // namespace oracle
