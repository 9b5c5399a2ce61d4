// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mpc_oracle.hpp).
// extern "C" surface of the CPU oracle so tests/ and bench.py's cpu_baseline leg
// can drive it through ctypes.  Consumes the SAME fp32 records the engine does
// (include/mpc_b200.h), widened to double, so parity compares arithmetic and
// not input rounding.  PARITY UNPINNED -- see the header of mpc_oracle.hpp.
#include "oracle_capi.h"

#include <omp.h>

#include <chrono>

#include "mpc_oracle.hpp"
#include "prep_oracle.hpp"

using namespace oracle;

namespace {

MpcParams to_params(const MpcConfig* c) {
  MpcParams p;
  p.H = c->horizon;
  p.dt = c->dt;
  p.mu = c->mu;
  p.fz_min = c->fz_min;
  p.fz_max = c->fz_max;
  p.mass = c->mass;
  for (int i = 0; i < 9; ++i) p.inertia[i] = c->inertia[i];
  for (int i = 0; i < 13; ++i) p.q_weights[i] = c->q_weights[i];
  for (int i = 0; i < 12; ++i) p.r_weights[i] = c->r_weights[i];
  const MpcSolverSettings& s = c->osqp;
  p.osqp.rho = s.rho; p.osqp.sigma = s.sigma; p.osqp.alpha = s.alpha;
  p.osqp.eps_abs = s.eps_abs; p.osqp.eps_rel = s.eps_rel;
  p.osqp.eps_prim_inf = s.eps_prim_inf; p.osqp.eps_dual_inf = s.eps_dual_inf;
  p.osqp.max_iter = s.max_iter; p.osqp.check_termination = s.check_termination;
  p.osqp.scaling = s.scaling; p.osqp.adaptive_rho = s.adaptive_rho;
  p.osqp.adaptive_rho_interval = s.adaptive_rho_interval;
  p.osqp.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  p.exact_discretization = c->exact_discretization != 0;
  p.foot_drift = c->foot_drift != 0;
  p.gait_aware = c->gait_aware != 0;
  return p;
}

BalanceParams to_params(const BalanceConfig* c) {
  BalanceParams p;
  for (int i = 0; i < 6; ++i) p.Q[i] = c->Q[i];
  p.R = c->R; p.mu = c->mu; p.F_min = c->F_min; p.F_max = c->F_max; p.mass = c->mass;
  for (int i = 0; i < 3; ++i) {
    p.kp_linear[i] = c->kp_linear[i]; p.kd_linear[i] = c->kd_linear[i];
    p.kp_angular[i] = c->kp_angular[i]; p.kd_angular[i] = c->kd_angular[i];
  }
  const MpcSolverSettings& s = c->osqp;
  p.osqp.rho = s.rho; p.osqp.sigma = s.sigma; p.osqp.alpha = s.alpha;
  p.osqp.eps_abs = s.eps_abs; p.osqp.eps_rel = s.eps_rel;
  p.osqp.eps_prim_inf = s.eps_prim_inf; p.osqp.eps_dual_inf = s.eps_dual_inf;
  p.osqp.max_iter = s.max_iter; p.osqp.check_termination = s.check_termination;
  p.osqp.scaling = s.scaling; p.osqp.adaptive_rho = s.adaptive_rho;
  p.osqp.adaptive_rho_interval = s.adaptive_rho_interval;
  p.osqp.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  return p;
}

template <class T>
RobotState<T> widen(const MpcStateIn& r) {
  RobotState<T> s;
  for (int i = 0; i < 3; ++i) {
    s.euler[i] = T(r.euler[i]); s.pos[i] = T(r.pos[i]);
    s.ang_vel[i] = T(r.ang_vel[i]); s.lin_vel[i] = T(r.lin_vel[i]);
    s.euler_d[i] = T(r.euler_d[i]); s.lin_vel_d[i] = T(r.lin_vel_d[i]);
    s.ang_vel_d[i] = T(r.ang_vel_d[i]);
  }
  s.pos_d_z = T(r.pos_d_z);
  for (int i = 0; i < 9; ++i) s.rot_mat[i] = T(r.rot_mat[i]);
  for (int i = 0; i < 12; ++i) s.foot_pos_abs[i] = T(r.foot_pos_abs[i]);
  for (int i = 0; i < 4; ++i) s.contacts[i] = r.contacts[i] != 0.0f;
  return s;
}

template <class T>
BalanceState<T> widen(const BalanceStateIn& r) {
  BalanceState<T> s;
  for (int i = 0; i < 3; ++i) {
    s.euler[i] = T(r.euler[i]); s.pos[i] = T(r.pos[i]);
    s.ang_vel[i] = T(r.ang_vel[i]); s.lin_vel[i] = T(r.lin_vel[i]);
    s.euler_d[i] = T(r.euler_d[i]); s.pos_d[i] = T(r.pos_d[i]);
    s.lin_vel_d[i] = T(r.lin_vel_d[i]); s.ang_vel_d[i] = T(r.ang_vel_d[i]);
  }
  for (int i = 0; i < 9; ++i) { s.rot_mat[i] = T(r.rot_mat[i]); s.rot_mat_z[i] = T(r.rot_mat_z[i]); }
  for (int i = 0; i < 12; ++i) s.foot_pos_abs[i] = T(r.foot_pos_abs[i]);
  for (int i = 0; i < 4; ++i) s.contacts[i] = r.contacts[i] != 0.0f;
  return s;
}

template <class T>
void to_result(const GrfResult<T>& g, OracleResult* o) {
  for (int i = 0; i < 12; ++i) o->grf[i] = double(g.grf[i]);
  o->status = g.status; o->iters = g.iters; o->rho_updates = g.rho_updates;
  o->pri_res = g.pri_res; o->dua_res = g.dua_res;
}

template <class T>
int grf_batch(const MpcConfig* cfg, const MpcStateIn* states, int n, OracleResult* out,
              double* solutions, int threads) {
  const MpcParams p = to_params(cfg);
  if (threads <= 0) threads = omp_get_max_threads();
  const int nvar = kNumDof * p.H;
#pragma omp parallel num_threads(threads)
  {
    MpcProblem<T> pb(p);
    std::vector<T> sol(nvar);
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
      RobotState<T> st = widen<T>(states[i]);
      GrfResult<T> g;
      mpc_build(p, st, pb);
      mpc_solve(p, st, pb, g, sol.data());
      to_result(g, &out[i]);
      if (solutions)
        for (int k = 0; k < nvar; ++k) solutions[size_t(i) * nvar + k] = double(sol[k]);
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int oracle_mpc_build_qp(const MpcConfig* cfg, const MpcStateIn* state, double* P, double* q,
                        double* l, double* u) {
  const MpcParams p = to_params(cfg);
  MpcProblem<double> pb(p);
  RobotState<double> st = widen<double>(*state);
  mpc_build(p, st, pb);
  const auto& m = pb.mpc;
  if (P) std::copy(m.hessian.begin(), m.hessian.end(), P);
  if (q) std::copy(m.gradient.begin(), m.gradient.end(), q);
  if (l) std::copy(m.lb.begin(), m.lb.end(), l);
  if (u) std::copy(m.ub.begin(), m.ub.end(), u);
  return 0;
}

int oracle_mpc_build_intermediates(const MpcConfig* cfg, const MpcStateIn* state, double* A_d,
                                   double* B_d, double* A_qp, double* B_qp, double* x0,
                                   double* x_ref) {
  const MpcParams p = to_params(cfg);
  MpcProblem<double> pb(p);
  RobotState<double> st = widen<double>(*state);
  mpc_build(p, st, pb);
  const auto& m = pb.mpc;
  if (A_d) std::copy(m.A_mat_d, m.A_mat_d + 169, A_d);
  if (B_d) std::copy(m.B_mat_d, m.B_mat_d + 156, B_d);
  if (A_qp) std::copy(m.A_qp.begin(), m.A_qp.end(), A_qp);
  if (B_qp) std::copy(m.B_qp.begin(), m.B_qp.end(), B_qp);
  if (x0) std::copy(pb.x0.begin(), pb.x0.end(), x0);
  if (x_ref) std::copy(pb.x_ref.begin(), pb.x_ref.end(), x_ref);
  return 0;
}

int oracle_mpc_compute_grf(const MpcConfig* cfg, const MpcStateIn* states, int32_t n,
                           OracleResult* out, double* solutions, int32_t threads) {
  return grf_batch<double>(cfg, states, n, out, solutions, threads);
}

int oracle_mpc_compute_grf_f32(const MpcConfig* cfg, const MpcStateIn* states, int32_t n,
                               OracleResult* out, double* solutions, int32_t threads) {
  return grf_batch<float>(cfg, states, n, out, solutions, threads);
}

int oracle_qp_mats_from_model(const MpcConfig* cfg, const double* A_mat_d,
                              const double* B_mat_d_list, const double* mpc_states,
                              const double* mpc_states_d, const int32_t* contacts,
                              double* hessian, double* gradient, double* lb, double* ub) {
  ConvexMpc<double> mpc(cfg->q_weights, cfg->r_weights, cfg->horizon);
  mpc.mu = cfg->mu;
  mpc.build_constraints();
  std::copy(A_mat_d, A_mat_d + 169, mpc.A_mat_d);
  std::copy(B_mat_d_list, B_mat_d_list + size_t(mpc.s) * kNumDof, mpc.B_mat_d_list.begin());
  bool c[4];
  for (int i = 0; i < 4; ++i) c[i] = contacts[i] != 0;
  mpc.calculate_qp_mats(mpc_states, mpc_states_d, c);
  if (hessian) std::copy(mpc.hessian.begin(), mpc.hessian.end(), hessian);
  if (gradient) std::copy(mpc.gradient.begin(), mpc.gradient.end(), gradient);
  if (lb) std::copy(mpc.lb.begin(), mpc.lb.end(), lb);
  if (ub) std::copy(mpc.ub.begin(), mpc.ub.end(), ub);
  return 0;
}

int oracle_osqp_solve_mpc(const MpcConfig* cfg, const double* P, const double* q, const double* l,
                          const double* u, double* x, double* y, OracleResult* info) {
  const MpcParams p = to_params(cfg);
  double qw[13] = {0}, rw[12] = {0};
  ConvexMpc<double> mpc(qw, rw, p.H);
  mpc.mu = p.mu;
  mpc.build_constraints();
  Csr<double> A;
  A.m = mpc.m; A.n = mpc.n;
  A.row_ptr = mpc.Ac_row_ptr; A.col = mpc.Ac_col; A.val = mpc.Ac_val;
  Osqp<double> solver;
  solver.setup(mpc.n, mpc.m, P, q, A, l, u, p.osqp);
  solver.solve();
  solver.solution(x, y);
  if (info) {
    for (int i = 0; i < 12; ++i) info->grf[i] = x[i];
    info->status = solver.info.status; info->iters = solver.info.iters;
    info->rho_updates = solver.info.rho_updates;
    info->pri_res = solver.info.pri_res; info->dua_res = solver.info.dua_res;
  }
  return 0;
}

// Experiment hook: the same solve instantiated in float on caller-supplied data.
int oracle_osqp_solve_mpc_f32(const MpcConfig* cfg, const double* P, const double* q,
                              const double* l, const double* u, double* x, OracleResult* info) {
  const MpcParams p = to_params(cfg);
  float qw[13] = {0}, rw[12] = {0};
  ConvexMpc<float> mpc(qw, rw, p.H);
  mpc.mu = float(p.mu);
  mpc.build_constraints();
  Csr<float> A;
  A.m = mpc.m; A.n = mpc.n;
  A.row_ptr = mpc.Ac_row_ptr; A.col = mpc.Ac_col; A.val = mpc.Ac_val;
  std::vector<float> Pf(size_t(mpc.n) * mpc.n), qf(mpc.n), lf(mpc.m), uf(mpc.m), xf(mpc.n);
  for (size_t i = 0; i < Pf.size(); ++i) Pf[i] = float(P[i]);
  for (int i = 0; i < mpc.n; ++i) qf[i] = float(q[i]);
  for (int i = 0; i < mpc.m; ++i) { lf[i] = float(l[i]); uf[i] = float(u[i]); }
  Osqp<float> solver;
  solver.setup(mpc.n, mpc.m, Pf.data(), qf.data(), A, lf.data(), uf.data(), p.osqp);
  solver.solve();
  solver.solution(xf.data(), nullptr);
  for (int i = 0; i < mpc.n; ++i) x[i] = xf[i];
  if (info) {
    info->status = solver.info.status; info->iters = solver.info.iters;
    info->rho_updates = solver.info.rho_updates;
    info->pri_res = solver.info.pri_res; info->dua_res = solver.info.dua_res;
  }
  return 0;
}

int oracle_balance_build_qp(const BalanceConfig* cfg, const BalanceStateIn* state, double* P,
                            double* q, double* l, double* u) {
  const BalanceParams p = to_params(cfg);
  BalanceState<double> st = widen<double>(*state);
  BalanceProblem<double> pb;
  balance_build(p, st, pb);
  if (P) std::copy(pb.P, pb.P + 144, P);
  if (q) std::copy(pb.q, pb.q + 12, q);
  if (l) std::copy(pb.l, pb.l + 20, l);
  if (u) std::copy(pb.u, pb.u + 20, u);
  return 0;
}

int oracle_balance_compute_grf(const BalanceConfig* cfg, const BalanceStateIn* states, int32_t n,
                               OracleResult* out, int32_t threads) {
  const BalanceParams p = to_params(cfg);
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    BalanceState<double> st = widen<double>(states[i]);
    BalanceProblem<double> pb;
    GrfResult<double> g;
    balance_build(p, st, pb);
    balance_solve(p, st, pb, g);
    to_result(g, &out[i]);
  }
  return 0;
}

// Warm-started streaming: `ticks` consecutive control ticks of n robots, states tick-major
// (states[t * n + i]); one persistent solver per robot, like the controller's member solver.
int oracle_mpc_stream(const MpcConfig* cfg, const MpcStateIn* states, int32_t n, int32_t ticks,
                      OracleResult* out, int32_t threads) {
  const MpcParams p = to_params(cfg);
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
  {
    MpcProblem<double> pb(p);
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
      MpcStream<double> stream;
      for (int t = 0; t < ticks; ++t) {
        RobotState<double> st = widen<double>(states[size_t(t) * n + i]);
        GrfResult<double> g;
        mpc_build(p, st, pb);
        stream.tick(p, st, pb, g);
        to_result(g, &out[size_t(t) * n + i]);
      }
    }
  }
  return 0;
}

// compute_joint_torques for n robots from given body-frame GRFs (n x 12 doubles); contacts are
// read at float word `contact_offset` of records `state_stride` floats apart (MPC 48/43,
// stance QP 64/54).
int oracle_torque_map(const float* state_words, int32_t state_stride, int32_t contact_offset,
                      const MpcTorqueIn* tin, const double* grf, int32_t n, double* joint_torques,
                      int32_t* nan_mask) {
  for (int i = 0; i < n; ++i) {
    double J[36], fk[12], km[3], tg[12];
    bool contacts[4];
    for (int k = 0; k < 36; ++k) J[k] = tin[i].j_foot[k];
    for (int k = 0; k < 12; ++k) { fk[k] = tin[i].foot_forces_kin[k]; tg[k] = tin[i].torques_gravity[k]; }
    for (int k = 0; k < 3; ++k) km[k] = tin[i].km_foot[k];
    for (int k = 0; k < 4; ++k) contacts[k] = state_words[size_t(i) * state_stride + contact_offset + k] != 0.0f;
    nan_mask[i] = torque_map<double>(J, contacts, grf + size_t(i) * 12, fk, km, tg, joint_torques + size_t(i) * 12);
  }
  return 0;
}

// State preparation for n robots over `ticks` consecutive sensor batches (tick-major), one
// persistent estimator / terrain filter per robot like the controller's members.
int oracle_prep_stream(const PrepConfig* cfg, const RobotSensorIn* sensors, int32_t n, int32_t ticks,
                       MpcStateIn* states, MpcTorqueIn* tin, RobotPrepOut* extras) {
  for (int i = 0; i < n; ++i) {
    prep_oracle::RobotSlot slot;
    for (int t = 0; t < ticks; ++t) {
      const size_t k = size_t(t) * n + i;
      prep_oracle::prep_tick(*cfg, sensors[k], slot, states[k], tin[k], extras[k]);
    }
  }
  return 0;
}

int oracle_leg_fk_jac(const double* rho_fix, const double* q, double* p, double* J) {
  prep_oracle::leg_fk_jac(rho_fix, q, p, J);
  return 0;
}

// the flags of cfg (exact_discretization, foot_drift, gait_aware) with the
// robots' gait records (may be NULL when gait_aware is 0).
int oracle_mpc_build_qp_ext(const MpcConfig* cfg, const MpcStateIn* state, const MpcGaitIn* gait, double* P, double* q,
                            double* l, double* u) {
  const MpcParams p = to_params(cfg);
  MpcProblem<double> pb(p);
  RobotState<double> st = widen<double>(*state);
  mpc_build(p, st, pb, gait);
  const int n = pb.mpc.n, m = pb.mpc.m;
  if (P) std::copy(pb.mpc.hessian.begin(), pb.mpc.hessian.begin() + size_t(n) * n, P);
  if (q) std::copy(pb.mpc.gradient.begin(), pb.mpc.gradient.begin() + n, q);
  if (l) std::copy(pb.mpc.lb.begin(), pb.mpc.lb.begin() + m, l);
  if (u) std::copy(pb.mpc.ub.begin(), pb.mpc.ub.begin() + m, u);
  return 0;
}

int oracle_mpc_compute_grf_ext(const MpcConfig* cfg, const MpcStateIn* states, const MpcGaitIn* gait, int32_t n,
                               OracleResult* out, int32_t threads) {
  const MpcParams p = to_params(cfg);
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
  {
    MpcProblem<double> pb(p);
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
      RobotState<double> st = widen<double>(states[i]);
      GrfResult<double> g;
      mpc_build(p, st, pb, gait ? gait + i : nullptr);
      mpc_solve(p, st, pb, g);
      to_result(g, &out[i]);
    }
  }
  return 0;
}

// the exact discretisation alone, for checking against a matrix exponential computed elsewhere
int oracle_discretize_exact(const MpcConfig* cfg, const MpcStateIn* state, double* A_d, double* B_d) {
  const MpcParams p = to_params(cfg);
  MpcProblem<double> pb(p);
  RobotState<double> st = widen<double>(*state);
  ConvexMpc<double>& mpc = pb.mpc;
  mpc.reset();
  mpc.calculate_A_mat_c(st.euler);
  double inertia[9];
  for (int i = 0; i < 9; ++i) inertia[i] = p.inertia[i];
  mpc.calculate_B_mat_c(p.mass, inertia, st.rot_mat, st.foot_pos_abs);
  mpc.state_space_discretization_exact(p.dt);
  std::copy(mpc.A_mat_d, mpc.A_mat_d + 169, A_d);
  std::copy(mpc.B_mat_d, mpc.B_mat_d + 156, B_d);
  return 0;
}

int oracle_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
