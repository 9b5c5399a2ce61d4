// C-ABI implementation of include/mpc_b200.h: engine life cycle, device memory,
// kernel launches.  No PyTorch, no CPU fallback: without a CUDA device every
// compute entry point fails with MPC_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mpc_b200.h"
#include "admm_kernel.cuh"
#include "balance_kernels.cuh"
#include "gen_kernels.cuh"
#include "mpc_kernels.cuh"
#include "prep_kernel.cuh"
#include "riccati_kernel.cuh"
#include "torque_map.cuh"
#include "wrench_kernel.cuh"
#include "wrench_riccati_kernel.cuh"
#include "wrench_tile_kernel.cuh"

using namespace mpcb200;

namespace {
thread_local std::string g_create_error;
}

struct MpcEngine {
  int kind = 0;  // 0: MPC branch, 1: stance-balance QP branch
  MpcConfig cfg{};
  BalanceConfig bcfg{};
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int capacity = 0;
  int n = 0;
  MpcStateIn* d_states_own = nullptr;
  const MpcStateIn* d_states = nullptr;
  BalanceStateIn* d_bstates = nullptr;
  double* d_P = nullptr;   // MPC: n x 120 x 128 f64 (padded rows)
  double* d_q = nullptr;   // MPC: n x 120 f64
  float* d_Pb = nullptr;   // balance QP parity read-back (f32)
  float* d_qb = nullptr;
  float* d_l = nullptr;
  float* d_u = nullptr;
  float* d_x = nullptr;
  MpcResult* d_results = nullptr;
  int* d_counter = nullptr;
  double* d_workspace = nullptr;  // long-horizon path: per-CTA B_qp / -K^-1 scratch
  int H = kH;                     // horizon
  int nvar() const { return 12 * H; }
  // doubles per warm-start slot: x | q | z | y | rho | live, padded (648 for H = 10, the layout of admm_kernel.cuh)
  int warm_stride() const { return H == kH ? kWarmStride : ((2 * nvar() + 2 * ncon() + 2 + 7) / 8) * 8; }
  int warm_live_offset() const { return 2 * nvar() + 2 * ncon() + 1; }
  int ncon() const { return 20 * H; }
  size_t p_stride() const { return H == kH ? size_t(kN) * kNP : size_t(nvar()) * nvar(); }  // doubles per problem
  MpcTorqueIn* d_tin = nullptr;   // optional torque-map inputs (compute_joint_torques), n records
  MpcTorqueOut* d_tout = nullptr;
  int torque_capacity = 0;
  bool torque_on = false;
  double* d_model = nullptr;      // structured solver: A_d | B_d list per problem (169 + 156 H doubles)
  bool structured = false;        // solve through the Riccati recursion instead of the dense inverse
  bool wrench = false;            // H = 10 default: fused build + wrench-space ADMM (wrench_kernel.cuh), no Hessian in HBM
  bool dense_ready = false;       // wrench engines: the dense QP (P, q, l, u) of the loaded states has been built for mpc_get_qp
  int dense_capacity = 0;         // problems the dense QP buffers can hold
  MpcGaitIn* d_gait = nullptr;    // gait scheduler records of the loaded states (gait_aware engines)
  int gait_capacity = 0;
  bool gait_on = false;
  RobotSensorIn* d_sensors = nullptr;  // state preparation: inputs, derived quantities, per-robot
  RobotPrepOut* d_extras = nullptr;    // estimator / terrain-filter slots (kPrepSlotStride doubles)
  double* d_prep_slots = nullptr;
  int prep_capacity = 0, prep_slot_capacity = 0;
  bool prepared = false;
  double* d_warm = nullptr;       // kWarmStride doubles per robot: the solver kept alive between ticks
  int warm_capacity = 0;
  long long* d_phase_clk = nullptr;  // optional per-phase cycle counters (mpc_debug_phase_cycles)
  bool built = false, solved = false;
  int64_t launches = 0;
  BuildParams bp{};
  SolveParams sp{};
  BalanceParams bal{};
  std::string err;
};

namespace {

int fail(MpcEngine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  else g_create_error = msg;
  return code;
}

#define CUDA_TRY(e, expr)                                                              \
  do {                                                                                 \
    cudaError_t _rc = (expr);                                                          \
    if (_rc != cudaSuccess)                                                            \
      return fail((e), MPC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_rc)); \
  } while (0)

SolveParams make_solve_params(const MpcSolverSettings& s, double mu) {
  SolveParams sp{};
  sp.rho = s.rho;
  sp.sigma = s.sigma;
  sp.alpha = s.alpha;
  sp.eps_abs = s.eps_abs;
  sp.eps_rel = s.eps_rel;
  sp.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  sp.mu = mu;
  sp.max_iter = s.max_iter;
  sp.check_termination = s.check_termination;
  sp.scaling = s.scaling;
  sp.adaptive_rho = s.adaptive_rho;
  sp.adaptive_rho_interval = s.adaptive_rho_interval;
  return sp;
}

void fill_model(MpcEngine* e, const MpcConfig* cfg) {
  e->bp.dt = cfg->dt;
  e->bp.mu = cfg->mu;
  e->bp.fz_min = cfg->fz_min;
  e->bp.fz_max = cfg->fz_max;
  e->bp.mass = cfg->mass;
  for (int i = 0; i < 9; ++i) e->bp.inertia[i] = cfg->inertia[i];
  for (int i = 0; i < 13; ++i) e->bp.Qd[i] = 2.0 * cfg->q_weights[i];  // ConvexMpc.cpp:20
  for (int i = 0; i < 12; ++i) e->bp.Rd[i] = 2.0 * cfg->r_weights[i];  // ConvexMpc.cpp:41
  e->bp.exact_discretization = cfg->exact_discretization != 0;
  e->bp.foot_drift = cfg->foot_drift != 0;
  e->bp.gait_aware = cfg->gait_aware != 0;
}

int validate_settings(const MpcSolverSettings& s, std::string* why) {
  if (!(s.rho > 0) || !(s.sigma > 0) || !(s.alpha > 0 && s.alpha < 2)) { *why = "rho/sigma/alpha out of range"; return -1; }
  if (s.max_iter <= 0 || s.check_termination < 0 || s.scaling < 0) { *why = "max_iter/check_termination/scaling out of range"; return -1; }
  if (s.adaptive_rho && s.adaptive_rho_interval <= 0) { *why = "adaptive_rho needs a pinned adaptive_rho_interval > 0"; return -1; }
  if (!(s.eps_abs >= 0) || !(s.eps_rel >= 0) || (s.eps_abs == 0 && s.eps_rel == 0)) { *why = "eps_abs/eps_rel out of range"; return -1; }
  return 0;
}

int open_device(int device, int* num_sms, std::string* why) {
  int count = 0;
  cudaError_t rc = cudaGetDeviceCount(&count);
  if (rc != cudaSuccess || count == 0) {
    *why = std::string("no CUDA device (") + (rc == cudaSuccess ? "count 0" : cudaGetErrorString(rc)) +
           "); this engine has no CPU fallback";
    (void)cudaGetLastError();
    return MPC_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) { *why = "device ordinal out of range"; return MPC_ERR_INVALID; }
  rc = cudaSetDevice(device);
  if (rc != cudaSuccess) { *why = cudaGetErrorString(rc); return MPC_ERR_CUDA; }
  cudaDeviceProp prop;
  rc = cudaGetDeviceProperties(&prop, device);
  if (rc != cudaSuccess) { *why = cudaGetErrorString(rc); return MPC_ERR_CUDA; }
  if (prop.major < 10) {
    *why = "device is not sm_100 class; kernels are built for sm_100a only";
    return MPC_ERR_UNSUPPORTED;
  }
  *num_sms = prop.multiProcessorCount;
  return MPC_OK;
}

void free_buffers(MpcEngine* e) {
  cudaFree(e->d_states_own);
  cudaFree(e->d_bstates);
  cudaFree(e->d_P);
  cudaFree(e->d_q);
  cudaFree(e->d_model);
  e->d_model = nullptr;
  cudaFree(e->d_Pb);
  cudaFree(e->d_qb);
  cudaFree(e->d_l);
  cudaFree(e->d_u);
  cudaFree(e->d_x);
  cudaFree(e->d_results);
  e->d_states_own = nullptr;
  e->d_bstates = nullptr;
  e->d_P = e->d_q = nullptr;
  e->d_Pb = e->d_qb = e->d_l = e->d_u = e->d_x = nullptr;
  e->d_results = nullptr;
  e->capacity = 0;
}

int reserve(MpcEngine* e, int n) {
  if (n <= e->capacity) return MPC_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  free_buffers(e);
  int cap = 1;
  while (cap < n) cap <<= 1;
  if (cap < 64) cap = 64;
  if (e->kind == 0 && e->H != kH && cap > n) cap = n < 64 ? 64 : n;  // 1 MB per problem: no slack
  if (e->kind == 0) {
    CUDA_TRY(e, cudaMalloc(&e->d_states_own, size_t(cap) * sizeof(MpcStateIn)));
    e->dense_capacity = 0;
    if (!e->wrench) {
      // the wrench-space engine never writes the QP to memory; mpc_get_qp builds it on demand
      CUDA_TRY(e, cudaMalloc(&e->d_P, size_t(cap) * e->p_stride() * sizeof(double)));
      CUDA_TRY(e, cudaMalloc(&e->d_q, size_t(cap) * e->nvar() * sizeof(double)));
      CUDA_TRY(e, cudaMalloc(&e->d_l, size_t(cap) * e->ncon() * sizeof(float)));
      CUDA_TRY(e, cudaMalloc(&e->d_u, size_t(cap) * e->ncon() * sizeof(float)));
      e->dense_capacity = cap;
    }
    CUDA_TRY(e, cudaMalloc(&e->d_x, size_t(cap) * e->nvar() * sizeof(float)));
    if (e->structured) CUDA_TRY(e, cudaMalloc(&e->d_model, size_t(cap) * (169 + 156 * e->H) * sizeof(double)));
  } else {
    CUDA_TRY(e, cudaMalloc(&e->d_bstates, size_t(cap) * sizeof(BalanceStateIn)));
    CUDA_TRY(e, cudaMalloc(&e->d_Pb, size_t(cap) * 144 * sizeof(float)));
    CUDA_TRY(e, cudaMalloc(&e->d_qb, size_t(cap) * 12 * sizeof(float)));
    CUDA_TRY(e, cudaMalloc(&e->d_l, size_t(cap) * 20 * sizeof(float)));
    CUDA_TRY(e, cudaMalloc(&e->d_u, size_t(cap) * 20 * sizeof(float)));
  }
  CUDA_TRY(e, cudaMalloc(&e->d_results, size_t(cap) * sizeof(MpcResult)));
  e->capacity = cap;
  return MPC_OK;
}

int launch_build(MpcEngine* e, const MpcStateIn* d_states, ModelIn model, int n, double* P, double* q,
                 float* l, float* u, double* model_out = nullptr) {
  if (e->H == kH) {
    const int grid = n < e->num_sms * 8 ? n : e->num_sms * 8;
    qp_build_kernel<<<grid, kBuildThreads, sizeof(BuildSmem), e->stream>>>(d_states, d_states ? e->d_gait : nullptr, model, n,
                                                                      model_out, P, q, l, u, e->bp);
  } else {
    const int grid = n < e->num_sms ? n : e->num_sms;
    gen_build_kernel<30><<<grid, kGenBuildThreads, sizeof(GenBuildSmem<30>), e->stream>>>(
        d_states, d_states ? e->d_gait : nullptr, model, n, model_out, P, q, l, u, e->d_workspace, e->bp);
  }
  ++e->launches;
  CUDA_TRY(e, cudaGetLastError());
  return MPC_OK;
}

int launch_solve(MpcEngine* e, const double* P, const double* q, const float* l, const float* u,
                 const MpcStateIn* d_states, MpcResult* res, float* x, int n, double* warm = nullptr,
                 bool with_torque = false) {
  const MpcTorqueIn* tin = (with_torque && d_states) ? e->d_tin : nullptr;
  CUDA_TRY(e, cudaMemsetAsync(e->d_counter, 0, sizeof(int), e->stream));
  const int grid = n < e->num_sms ? n : e->num_sms;
  if (e->structured && (!warm || e->H != kH) && P == e->d_P) {
    // Riccati-structured ADMM (riccati_kernel.cuh): any horizon, one CTA of 128 threads per problem
    if (e->H == kH) {
      const int g = n < e->num_sms * 3 ? n : e->num_sms * 3;
      riccati_solve_kernel<kH><<<g, kRicThreads, sizeof(RicSmem<kH>), e->stream>>>(
          P, e->p_stride(), kNP, q, l, u, e->d_model, d_states, res, x, n, e->d_counter, nullptr, 0, e->bp, e->sp);
    } else {
      const int g = n < e->num_sms ? n : e->num_sms;
      riccati_solve_kernel<30><<<g, kRicThreads, sizeof(RicSmem<30>), e->stream>>>(
          P, e->p_stride(), e->nvar(), q, l, u, e->d_model, d_states, res, x, n, e->d_counter, warm, e->warm_stride(),
          e->bp, e->sp);
    }
    ++e->launches;
    CUDA_TRY(e, cudaGetLastError());
    if (tin) {
      torque_map_kernel<<<(4 * n + 127) / 128, 128, 0, e->stream>>>(
          res, reinterpret_cast<const float*>(d_states), int(sizeof(MpcStateIn) / 4), kOffContacts, tin, e->d_tout, n);
      ++e->launches;
      CUDA_TRY(e, cudaGetLastError());
    }
    return MPC_OK;
  }
  if (warm)
    admm_solve_kernel<false, true><<<grid, kSolveThreads, sizeof(SolveSmem), e->stream>>>(
        P, q, l, u, d_states, res, x, n, e->d_counter, nullptr, warm, tin, e->d_tout, e->sp);
  else if (e->H != kH) {
    gen_solve_kernel<30><<<grid, kGenSolveThreads, sizeof(GenSolveSmem<30>), e->stream>>>(
        P, q, l, u, d_states, res, x, n, e->d_counter, e->d_workspace, e->sp);
    if (tin) {
      // the long-horizon writer is not fused: map the written results
      torque_map_kernel<<<(4 * n + 127) / 128, 128, 0, e->stream>>>(
          res, reinterpret_cast<const float*>(d_states), int(sizeof(MpcStateIn) / 4), kOffContacts, tin, e->d_tout, n);
      ++e->launches;
    }
  }
  else if (e->d_phase_clk)
    admm_solve_kernel<true, false><<<grid, kSolveThreads, sizeof(SolveSmem), e->stream>>>(
        P, q, l, u, d_states, res, x, n, e->d_counter, e->d_phase_clk, nullptr, tin, e->d_tout, e->sp);
  else
    admm_solve_kernel<false, false><<<grid, kSolveThreads, sizeof(SolveSmem), e->stream>>>(
        P, q, l, u, d_states, res, x, n, e->d_counter, nullptr, nullptr, tin, e->d_tout, e->sp);
  ++e->launches;
  CUDA_TRY(e, cudaGetLastError());
  return MPC_OK;
}

// H = 10 default path: one fused kernel from the state records to the results
int launch_wrench(MpcEngine* e, int n, double* warm, bool with_torque) {
  const MpcTorqueIn* tin = with_torque ? e->d_tin : nullptr;
  CUDA_TRY(e, cudaMemsetAsync(e->d_counter, 0, sizeof(int), e->stream));
  if (e->H != kH) {
    const int full = e->num_sms * kWrcCtasPerSm;
    const int grid = n < full ? n : full;
    wrench_riccati_kernel<30><<<grid, kWrcThreads, sizeof(WrcSmem<30>), e->stream>>>(
        e->d_states, e->d_gait, e->d_results, e->d_x, n, e->d_counter, warm, e->warm_stride(), e->d_workspace, tin, e->d_tout,
        e->bp, e->sp);
    ++e->launches;
    CUDA_TRY(e, cudaGetLastError());
    return MPC_OK;
  }
  const int full = e->num_sms * kWrCtasPerSm;
  const int grid = n < full ? n : full;
  // the 2-D tiled core is the default; MPC_WRENCH_TILE=0 (read once per process) selects the half-row kernel
  static const bool tile = [] { const char* v = std::getenv("MPC_WRENCH_TILE"); return !(v && v[0] == '0'); }();
  if (tile)
    wrench_tile_kernel<kWrCtasPerSm><<<grid, kWrThreads, sizeof(WrenchSmem), e->stream>>>(
        e->d_states, e->d_gait, e->d_results, e->d_x, n, e->d_counter, warm, tin, e->d_tout, e->bp, e->sp);
  else
  wrench_solve_kernel<kWrCtasPerSm><<<grid, kWrThreads, sizeof(WrenchSmem), e->stream>>>(
      e->d_states, e->d_gait, e->d_results, e->d_x, n, e->d_counter, warm, tin, e->d_tout, e->bp, e->sp);
  ++e->launches;
  CUDA_TRY(e, cudaGetLastError());
  return MPC_OK;
}

// wrench engines: dense QP buffers for the parity / debug read-back
int ensure_dense(MpcEngine* e) {
  if (e->dense_capacity >= e->n && e->d_P) return MPC_OK;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  cudaFree(e->d_P); cudaFree(e->d_q); cudaFree(e->d_l); cudaFree(e->d_u);
  e->d_P = e->d_q = nullptr;
  e->d_l = e->d_u = nullptr;
  e->dense_capacity = 0;
  const size_t cap = size_t(e->n < 64 ? 64 : e->n);
  CUDA_TRY(e, cudaMalloc(&e->d_P, cap * e->p_stride() * sizeof(double)));
  CUDA_TRY(e, cudaMalloc(&e->d_q, cap * e->nvar() * sizeof(double)));
  CUDA_TRY(e, cudaMalloc(&e->d_l, cap * e->ncon() * sizeof(float)));
  CUDA_TRY(e, cudaMalloc(&e->d_u, cap * e->ncon() * sizeof(float)));
  e->dense_capacity = int(cap);
  return MPC_OK;
}

int reserve_torque(MpcEngine* e, int n) {
  if (n <= e->torque_capacity) return MPC_OK;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  cudaFree(e->d_tin);
  cudaFree(e->d_tout);
  e->d_tin = nullptr;
  e->d_tout = nullptr;
  e->torque_capacity = 0;
  CUDA_TRY(e, cudaMalloc(&e->d_tin, size_t(n) * sizeof(MpcTorqueIn)));
  CUDA_TRY(e, cudaMalloc(&e->d_tout, size_t(n) * sizeof(MpcTorqueOut)));
  e->torque_capacity = n;
  return MPC_OK;
}

int create_common(int kind, int device, MpcEngine** out) {
  if (!out) return fail(nullptr, MPC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  std::string why;
  int num_sms = 0;
  int rc = open_device(device, &num_sms, &why);
  if (rc != MPC_OK) return fail(nullptr, rc, why);
  MpcEngine* e = new (std::nothrow) MpcEngine();
  if (!e) return fail(nullptr, MPC_ERR_INVALID, "out of host memory");
  e->kind = kind;
  e->device = device;
  e->num_sms = num_sms;
  cudaError_t crc = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
  if (crc == cudaSuccess) crc = cudaMalloc(&e->d_counter, sizeof(int));
  const char* what = "stream / counter";
  auto opt_in = [&](const void* fn, size_t bytes, const char* name) {
    if (crc != cudaSuccess) return;
    crc = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (crc != cudaSuccess) what = name;
  };
  opt_in((const void*)qp_build_kernel, sizeof(BuildSmem), "qp_build_kernel shared memory");
  opt_in((const void*)admm_solve_kernel<false, false>, sizeof(SolveSmem), "admm_solve_kernel shared memory");
  opt_in((const void*)admm_solve_kernel<true, false>, sizeof(SolveSmem), "admm_solve_kernel (profile) shared memory");
  opt_in((const void*)admm_solve_kernel<false, true>, sizeof(SolveSmem), "admm_solve_kernel (warm) shared memory");
  opt_in((const void*)riccati_solve_kernel<kH>, sizeof(RicSmem<kH>), "riccati_solve_kernel<10> shared memory");
  opt_in((const void*)riccati_solve_kernel<30>, sizeof(RicSmem<30>), "riccati_solve_kernel<30> shared memory");
  opt_in((const void*)gen_build_kernel<30>, sizeof(GenBuildSmem<30>), "gen_build_kernel<30> shared memory");
  opt_in((const void*)wrench_riccati_kernel<30>, sizeof(WrcSmem<30>), "wrench_riccati_kernel<30> shared memory");
  opt_in((const void*)gen_solve_kernel<30>, sizeof(GenSolveSmem<30>), "gen_solve_kernel<30> shared memory");
  opt_in((const void*)wrench_solve_kernel<kWrCtasPerSm>, sizeof(WrenchSmem), "wrench_solve_kernel shared memory");
  opt_in((const void*)wrench_tile_kernel<kWrCtasPerSm>, sizeof(WrenchSmem), "wrench_tile_kernel shared memory");
  if (crc != cudaSuccess) {
    std::string msg = std::string("engine setup (") + what + "): " + cudaGetErrorString(crc);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    cudaFree(e->d_counter);
    delete e;
    return fail(nullptr, MPC_ERR_CUDA, msg);
  }
  e->stream = e->own_stream;
  *out = e;
  return MPC_OK;
}

}  // namespace

// register-only DFMA chains, eight independent accumulators per thread (scripts/fp64_bench.cu)
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = double(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" {

int mpc_measure_fp64_peak(int32_t device, double ms_target, double* tflops) {
  if (!tflops) return MPC_ERR_INVALID;
  std::string why;
  int num_sms = 0;
  int rc = open_device(device, &num_sms, &why);
  if (rc != MPC_OK) return fail(nullptr, rc, why);
  const int blocks = num_sms * 4, threads = 512;
  double* out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaMalloc(&out, size_t(blocks) * threads * sizeof(double)) != cudaSuccess ||
      cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
    cudaFree(out);
    return fail(nullptr, MPC_ERR_CUDA, "fp64 peak probe: allocation failed");
  }
  int iters = 2000;
  float ms = 0.f;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = MPC_ERR_CUDA; break; }
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * double(blocks) * threads * double(iters) * 8.0 / (double(ms) * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
    // size the next launch for the requested duration
    if (ms > 0.f) iters = int(double(iters) * (ms_target > 0 ? ms_target : 3.0) / double(ms));
    if (iters < 1000) iters = 1000;
    if (iters > 4000000) iters = 4000000;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (rc != MPC_OK) return fail(nullptr, rc, "fp64 peak probe failed");
  *tflops = best;
  return MPC_OK;
}

int mpc_engine_create(const MpcConfig* cfg, int32_t device, MpcEngine** out) {
  if (!cfg) return fail(nullptr, MPC_ERR_INVALID, "cfg is NULL");
  if (cfg->horizon != kH && cfg->horizon != 30)
    return fail(nullptr, MPC_ERR_UNSUPPORTED,
                "horizon " + std::to_string(cfg->horizon) + " not built; this build has H = 10 and H = 30 kernels");
  std::string why;
  if (validate_settings(cfg->osqp, &why)) return fail(nullptr, MPC_ERR_INVALID, why);
  if (!(cfg->dt > 0) || !(cfg->mass > 0) || !(cfg->mu > 0))
    return fail(nullptr, MPC_ERR_INVALID, "dt/mass/mu must be positive");
  MpcEngine* e = nullptr;
  int rc = create_common(0, device, &e);
  if (rc != MPC_OK) return rc;
  e->cfg = *cfg;
  e->H = cfg->horizon;
  if (e->H != kH) {
    // long-horizon path: one scratch matrix per persistent CTA (B_qp is 13H x 12H, -K^-1 is 12H x 12H)
    const size_t per_cta = size_t(13 * e->H) * (12 * e->H);
    cudaError_t wrc = cudaMalloc(&e->d_workspace, size_t(e->num_sms) * per_cta * sizeof(double));
    if (wrc != cudaSuccess) {
      const std::string msg = std::string("workspace: ") + cudaGetErrorString(wrc);
      mpc_engine_destroy(e);
      return fail(nullptr, MPC_ERR_CUDA, msg);
    }
  }
  fill_model(e, cfg);
  if (cfg->structured_solver < 0 || cfg->structured_solver > 3 ||
      (cfg->structured_solver == 3 && cfg->horizon != kH && cfg->exact_discretization)) {
    mpc_engine_destroy(e);
    return fail(nullptr, MPC_ERR_INVALID,
                "structured_solver must be 0..3 (3 = wrench-space; at horizon 30 not with exact_discretization)");
  }
  // wrench-space engines (fused build + solve, nothing in HBM): H = 10 dense 60 x 60 core, H = 30 six-input Riccati
  e->wrench = (cfg->structured_solver == 0 || cfg->structured_solver == 3) &&
              (cfg->horizon == kH || !cfg->exact_discretization);
  e->structured = cfg->structured_solver == 1 || (cfg->structured_solver == 0 && cfg->horizon != kH && !e->wrench);
  e->sp = make_solve_params(cfg->osqp, cfg->mu);
  *out = e;
  return MPC_OK;
}

int balance_engine_create(const BalanceConfig* cfg, int32_t device, MpcEngine** out) {
  if (!cfg) return fail(nullptr, MPC_ERR_INVALID, "cfg is NULL");
  std::string why;
  if (validate_settings(cfg->osqp, &why)) return fail(nullptr, MPC_ERR_INVALID, why);
  MpcEngine* e = nullptr;
  int rc = create_common(1, device, &e);
  if (rc != MPC_OK) return rc;
  e->bcfg = *cfg;
  e->bal = make_balance_params(*cfg);
  *out = e;
  return MPC_OK;
}

void mpc_engine_destroy(MpcEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  free_buffers(e);
  cudaFree(e->d_counter);
  cudaFree(e->d_phase_clk);
  cudaFree(e->d_workspace);
  cudaFree(e->d_warm);
  cudaFree(e->d_tin);
  cudaFree(e->d_tout);
  cudaFree(e->d_sensors);
  cudaFree(e->d_extras);
  cudaFree(e->d_prep_slots);
  cudaFree(e->d_gait);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

const char* mpc_last_error(const MpcEngine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int mpc_set_stream(MpcEngine* e, void* cuda_stream) {
  if (!e) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return MPC_OK;
}

int mpc_synchronize(MpcEngine* e) {
  if (!e) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return MPC_OK;
}

int64_t mpc_kernel_launches(const MpcEngine* e) { return e ? e->launches : 0; }

int mpc_debug_phase_cycles(MpcEngine* e, int32_t enable, int64_t* out6) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  const size_t n6 = size_t(e->num_sms) * 6;
  const size_t n = size_t(e->num_sms) * 22;  // 6 phase + 16 fine counters per CTA
  if (out6 && e->d_phase_clk) {
    std::vector<long long> h(n);
    CUDA_TRY(e, cudaMemcpy(h.data(), e->d_phase_clk, n * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 22; ++i) out6[i] = 0;
    for (size_t k = 0; k < n6; ++k) out6[k % 6] += h[k];
    for (size_t k = n6; k < n; ++k) out6[6 + (k - n6) % 16] += h[k];
  }
  if (enable && !e->d_phase_clk) {
    CUDA_TRY(e, cudaMalloc(&e->d_phase_clk, n * sizeof(long long)));
  } else if (!enable && e->d_phase_clk) {
    cudaFree(e->d_phase_clk);
    e->d_phase_clk = nullptr;
  }
  if (e->d_phase_clk) CUDA_TRY(e, cudaMemset(e->d_phase_clk, 0, n * sizeof(long long)));
  return MPC_OK;
}

int mpc_load_states(MpcEngine* e, const MpcStateIn* host, int32_t n) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (n < 0 || (n > 0 && !host)) return fail(e, MPC_ERR_INVALID, "bad state buffer");
  CUDA_TRY(e, cudaSetDevice(e->device));
  int rc = reserve(e, n);
  if (rc) return rc;
  if (n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(e->d_states_own, host, size_t(n) * sizeof(MpcStateIn),
                                cudaMemcpyHostToDevice, e->stream));
  e->d_states = e->d_states_own;
  e->n = n;
  e->built = e->solved = false;
  e->torque_on = false;  // torque and gait inputs belong to one batch of states
  e->gait_on = false;
  e->prepared = false;
  return MPC_OK;
}

int mpc_set_states_device(MpcEngine* e, const MpcStateIn* dev, int32_t n) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (n < 0 || (n > 0 && !dev)) return fail(e, MPC_ERR_INVALID, "bad device state buffer");
  CUDA_TRY(e, cudaSetDevice(e->device));
  int rc = reserve(e, n);
  if (rc) return rc;
  e->d_states = dev;
  e->n = n;
  e->built = e->solved = false;
  e->torque_on = false;  // torque and gait inputs belong to one batch of states
  e->gait_on = false;
  e->prepared = false;
  return MPC_OK;
}

int mpc_build_qp_async(MpcEngine* e) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!e->d_states && e->n > 0) return fail(e, MPC_ERR_STATE, "mpc_build_qp before mpc_load_states");
  if (e->bp.gait_aware && !e->gait_on && e->n > 0)
    return fail(e, MPC_ERR_STATE, "gait_aware engine: mpc_set_gait_inputs must follow the state load");
  CUDA_TRY(e, cudaSetDevice(e->device));
  e->dense_ready = false;
  if (e->n > 0 && !e->wrench) {
    // (wrench engines build inside the solve kernel; mpc_get_qp builds the dense QP on demand)
    ModelIn none{};
    int rc = launch_build(e, e->d_states, none, e->n, e->d_P, e->d_q, e->d_l, e->d_u, e->structured ? e->d_model : nullptr);
    if (rc) return rc;
  }
  e->built = true;
  e->solved = false;
  return MPC_OK;
}

int mpc_build_qp(MpcEngine* e) {
  int rc = mpc_build_qp_async(e);
  if (rc) return rc;
  return mpc_synchronize(e);
}

int mpc_get_qp(MpcEngine* e, int32_t idx, float* P, float* q, float* l, float* u) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!e->built) return fail(e, MPC_ERR_STATE, "mpc_get_qp before mpc_build_qp");
  if (idx < 0 || idx >= e->n) return fail(e, MPC_ERR_INVALID, "problem index out of range");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->wrench && !e->dense_ready) {
    int rc = ensure_dense(e);
    if (rc) return rc;
    ModelIn none{};
    rc = launch_build(e, e->d_states, none, e->n, e->d_P, e->d_q, e->d_l, e->d_u, nullptr);
    if (rc) return rc;
    e->dense_ready = true;
  }
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  const int nv = e->nvar(), nc = e->ncon();
  const size_t stride = e->H == kH ? size_t(kNP) : size_t(nv);
  if (P) {
    std::vector<double> hp(e->p_stride());
    CUDA_TRY(e, cudaMemcpy(hp.data(), e->d_P + size_t(idx) * e->p_stride(), hp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int r = 0; r < nv; ++r)
      for (int c = 0; c < nv; ++c) P[size_t(r) * nv + c] = (float)hp[size_t(r) * stride + c];
  }
  if (q) {
    std::vector<double> hq(nv);
    CUDA_TRY(e, cudaMemcpy(hq.data(), e->d_q + size_t(idx) * nv, nv * sizeof(double), cudaMemcpyDeviceToHost));
    for (int c = 0; c < nv; ++c) q[c] = (float)hq[c];
  }
  if (l) CUDA_TRY(e, cudaMemcpy(l, e->d_l + size_t(idx) * nc, nc * sizeof(float), cudaMemcpyDeviceToHost));
  if (u) CUDA_TRY(e, cudaMemcpy(u, e->d_u + size_t(idx) * nc, nc * sizeof(float), cudaMemcpyDeviceToHost));
  return MPC_OK;
}

int mpc_solve_async(MpcEngine* e) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!e->built) return fail(e, MPC_ERR_STATE, "mpc_solve before mpc_build_qp");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->n > 0) {
    int rc = e->wrench ? launch_wrench(e, e->n, nullptr, e->torque_on)
                       : launch_solve(e, e->d_P, e->d_q, e->d_l, e->d_u, e->d_states, e->d_results, e->d_x, e->n, nullptr,
                                      e->torque_on);
    if (rc) return rc;
  }
  e->solved = true;
  return MPC_OK;
}

int mpc_solve(MpcEngine* e) {
  int rc = mpc_solve_async(e);
  if (rc) return rc;
  return mpc_synchronize(e);
}

// ---- gait-aware horizon -----------------------------------------------------------

int mpc_set_gait_inputs(MpcEngine* e, const MpcGaitIn* host, int32_t n) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!host) {
    e->gait_on = false;
    return MPC_OK;
  }
  if (n != e->n || (!e->d_states && n > 0))
    return fail(e, MPC_ERR_STATE, "mpc_set_gait_inputs: load the n states first, then give n gait records");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (n > e->gait_capacity) {
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    cudaFree(e->d_gait);
    e->d_gait = nullptr;
    e->gait_capacity = 0;
    CUDA_TRY(e, cudaMalloc(&e->d_gait, size_t(n) * sizeof(MpcGaitIn)));
    e->gait_capacity = n;
  }
  if (n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(e->d_gait, host, size_t(n) * sizeof(MpcGaitIn), cudaMemcpyHostToDevice, e->stream));
  e->gait_on = true;
  e->built = e->solved = false;
  return MPC_OK;
}

// ---- torque map (compute_joint_torques) ----------------------------------------

int mpc_set_torque_inputs(MpcEngine* e, const MpcTorqueIn* host, int32_t n) {
  if (!e) return MPC_ERR_INVALID;
  if (!host) {
    e->torque_on = false;
    return MPC_OK;
  }
  if (n != e->n || (e->kind == 0 && !e->d_states && n > 0) || (e->kind == 1 && !e->built))
    return fail(e, MPC_ERR_STATE, "mpc_set_torque_inputs: load the n states first, then give n torque records");
  CUDA_TRY(e, cudaSetDevice(e->device));
  int rc = reserve_torque(e, n);
  if (rc) return rc;
  if (n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(e->d_tin, host, size_t(n) * sizeof(MpcTorqueIn), cudaMemcpyHostToDevice, e->stream));
  e->torque_on = true;
  e->solved = false;
  return MPC_OK;
}

int mpc_get_torques(MpcEngine* e, MpcTorqueOut* host) {
  if (!e) return MPC_ERR_INVALID;
  if (!e->solved || !e->torque_on) return fail(e, MPC_ERR_STATE, "mpc_get_torques: no solve with torque inputs yet");
  if (e->n > 0 && !host) return fail(e, MPC_ERR_INVALID, "host buffer is NULL");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(host, e->d_tout, size_t(e->n) * sizeof(MpcTorqueOut), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return MPC_OK;
}

// ---- upstream state preparation -------------------------------------------------

int mpc_prepare_states(MpcEngine* e, const PrepConfig* cfg, const RobotSensorIn* host, int32_t n) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!cfg || n < 0 || (n > 0 && !host)) return fail(e, MPC_ERR_INVALID, "bad sensor buffer or config");
  CUDA_TRY(e, cudaSetDevice(e->device));
  int rc = reserve(e, n);
  if (rc) return rc;
  rc = reserve_torque(e, n);
  if (rc) return rc;
  if (n > e->prep_capacity) {
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    cudaFree(e->d_sensors);
    cudaFree(e->d_extras);
    e->d_sensors = nullptr;
    e->d_extras = nullptr;
    e->prep_capacity = 0;
    CUDA_TRY(e, cudaMalloc(&e->d_sensors, size_t(n) * sizeof(RobotSensorIn)));
    CUDA_TRY(e, cudaMalloc(&e->d_extras, size_t(n) * sizeof(RobotPrepOut)));
    e->prep_capacity = n;
  }
  if (n > e->prep_slot_capacity) {
    // growing keeps the estimators of the robots that already have a slot
    double* grown = nullptr;
    const size_t bytes = size_t(n) * kPrepSlotStride * sizeof(double);
    CUDA_TRY(e, cudaMalloc(&grown, bytes));
    cudaError_t crc = cudaMemsetAsync(grown, 0, bytes, e->stream);
    if (crc == cudaSuccess && e->d_prep_slots)
      crc = cudaMemcpyAsync(grown, e->d_prep_slots, size_t(e->prep_slot_capacity) * kPrepSlotStride * sizeof(double),
                            cudaMemcpyDeviceToDevice, e->stream);
    if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
    if (crc != cudaSuccess) {
      cudaFree(grown);
      return fail(e, MPC_ERR_CUDA, cudaGetErrorString(crc));
    }
    cudaFree(e->d_prep_slots);
    e->d_prep_slots = grown;
    e->prep_slot_capacity = n;
  }
  if (n > 0) {
    CUDA_TRY(e, cudaMemcpyAsync(e->d_sensors, host, size_t(n) * sizeof(RobotSensorIn), cudaMemcpyHostToDevice, e->stream));
    PrepParams pp{};
    for (int i = 0; i < 20; ++i) pp.rho_fix[i] = cfg->rho_fix[i];
    for (int i = 0; i < 3; ++i) pp.km_foot[i] = cfg->km_foot[i];
    for (int i = 0; i < 12; ++i) pp.torques_gravity[i] = cfg->torques_gravity[i];
    pp.use_estimator = cfg->use_estimator;
    pp.assume_flat_ground = cfg->assume_flat_ground;
    pp.use_terrain_adapt = cfg->use_terrain_adapt;
    const int grid = (n + kPrepWarps - 1) / kPrepWarps;
    state_prep_kernel<<<grid, 32 * kPrepWarps, kPrepWarps * sizeof(PrepWarpSmem), e->stream>>>(
        e->d_sensors, n, e->d_prep_slots, e->d_states_own, e->d_tin, e->d_extras, pp);
    ++e->launches;
    CUDA_TRY(e, cudaGetLastError());
  }
  e->d_states = e->d_states_own;
  e->n = n;
  e->built = e->solved = false;
  e->torque_on = true;
  e->gait_on = false;
  e->prepared = true;
  return MPC_OK;
}

int mpc_get_prepared(MpcEngine* e, MpcStateIn* states, MpcTorqueIn* torque_in, RobotPrepOut* extras) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!e->prepared) return fail(e, MPC_ERR_STATE, "mpc_get_prepared before mpc_prepare_states");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const size_t n = size_t(e->n);
  if (n > 0 && states)
    CUDA_TRY(e, cudaMemcpyAsync(states, e->d_states_own, n * sizeof(MpcStateIn), cudaMemcpyDeviceToHost, e->stream));
  if (n > 0 && torque_in)
    CUDA_TRY(e, cudaMemcpyAsync(torque_in, e->d_tin, n * sizeof(MpcTorqueIn), cudaMemcpyDeviceToHost, e->stream));
  if (n > 0 && extras)
    CUDA_TRY(e, cudaMemcpyAsync(extras, e->d_extras, n * sizeof(RobotPrepOut), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return MPC_OK;
}

int mpc_prepare_reset(MpcEngine* e) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->d_prep_slots)
    CUDA_TRY(e, cudaMemsetAsync(e->d_prep_slots, 0, size_t(e->prep_slot_capacity) * kPrepSlotStride * sizeof(double), e->stream));
  return MPC_OK;
}

// ---- warm-started streaming (one persistent solver per robot slot) ------------

int mpc_stream_reset(MpcEngine* e) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->d_warm)
    CUDA_TRY(e, cudaMemsetAsync(e->d_warm, 0, size_t(e->warm_capacity) * e->warm_stride() * sizeof(double), e->stream));
  return MPC_OK;
}

int mpc_stream_reset_slots(MpcEngine* e, const int32_t* idx, int32_t k) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (k < 0 || (k > 0 && !idx)) return fail(e, MPC_ERR_INVALID, "bad slot list");
  CUDA_TRY(e, cudaSetDevice(e->device));
  for (int i = 0; i < k; ++i) {
    if (idx[i] < 0) return fail(e, MPC_ERR_INVALID, "negative slot index");
    if (idx[i] >= e->warm_capacity) continue;  // a slot that never solved is not live
    CUDA_TRY(e, cudaMemsetAsync(e->d_warm + size_t(idx[i]) * e->warm_stride() + e->warm_live_offset(), 0, sizeof(double),
                                e->stream));
  }
  return MPC_OK;
}

int mpc_engine_update_model(MpcEngine* e, const MpcConfig* cfg) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!cfg) return fail(e, MPC_ERR_INVALID, "cfg is NULL");
  if (cfg->horizon != e->cfg.horizon || cfg->structured_solver != e->cfg.structured_solver ||
      (cfg->exact_discretization != 0) != (e->cfg.exact_discretization != 0) ||
      (cfg->foot_drift != 0) != (e->cfg.foot_drift != 0) || (cfg->gait_aware != 0) != (e->cfg.gait_aware != 0))
    return fail(e, MPC_ERR_INVALID, "mpc_engine_update_model: horizon, extension flags and structured_solver are fixed at creation");
  std::string why;
  if (validate_settings(cfg->osqp, &why)) return fail(e, MPC_ERR_INVALID, why);
  if (!(cfg->dt > 0) || !(cfg->mass > 0) || !(cfg->mu > 0)) return fail(e, MPC_ERR_INVALID, "dt/mass/mu must be positive");
  e->cfg = *cfg;
  fill_model(e, cfg);
  e->sp = make_solve_params(cfg->osqp, cfg->mu);
  e->built = e->solved = false;  // the loaded states must be rebuilt with the new constants
  e->dense_ready = false;
  return MPC_OK;
}

int mpc_solve_warm_async(MpcEngine* e) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (e->H != kH && !e->structured && !e->wrench)
    return fail(e, MPC_ERR_UNSUPPORTED, "warm-started long-horizon solves need a structured solver (structured_solver 0, 1 or 3)");
  if (!e->built) return fail(e, MPC_ERR_STATE, "mpc_solve_warm before mpc_build_qp");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->n > e->warm_capacity) {
    // growing the slot array keeps the live solvers of the slots that already exist
    double* grown = nullptr;
    const size_t ws = size_t(e->warm_stride());
    CUDA_TRY(e, cudaMalloc(&grown, size_t(e->n) * ws * sizeof(double)));
    cudaError_t crc = cudaMemsetAsync(grown, 0, size_t(e->n) * ws * sizeof(double), e->stream);
    if (crc == cudaSuccess && e->d_warm)
      crc = cudaMemcpyAsync(grown, e->d_warm, size_t(e->warm_capacity) * ws * sizeof(double),
                            cudaMemcpyDeviceToDevice, e->stream);
    if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
    if (crc != cudaSuccess) {
      cudaFree(grown);
      return fail(e, MPC_ERR_CUDA, cudaGetErrorString(crc));
    }
    cudaFree(e->d_warm);
    e->d_warm = grown;
    e->warm_capacity = e->n;
  }
  if (e->n > 0) {
    int rc = e->wrench ? launch_wrench(e, e->n, e->d_warm, e->torque_on)
                       : launch_solve(e, e->d_P, e->d_q, e->d_l, e->d_u, e->d_states, e->d_results, e->d_x, e->n, e->d_warm,
                                      e->torque_on);
    if (rc) return rc;
  }
  e->solved = true;
  return MPC_OK;
}

int mpc_solve_warm(MpcEngine* e) {
  int rc = mpc_solve_warm_async(e);
  if (rc) return rc;
  return mpc_synchronize(e);
}

int mpc_stream_step(MpcEngine* e, const MpcStateIn* host_in, MpcResult* host_out, int32_t n) {
  int rc = mpc_load_states(e, host_in, n);
  if (rc) return rc;
  rc = mpc_build_qp_async(e);
  if (rc) return rc;
  rc = mpc_solve_warm_async(e);
  if (rc) return rc;
  return mpc_get_results(e, host_out);
}

int mpc_get_results(MpcEngine* e, MpcResult* host) {
  if (!e) return MPC_ERR_INVALID;
  if (!e->solved) return fail(e, MPC_ERR_STATE, "mpc_get_results before mpc_solve");
  if (e->n > 0 && !host) return fail(e, MPC_ERR_INVALID, "host buffer is NULL");
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(host, e->d_results, size_t(e->n) * sizeof(MpcResult),
                                cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return MPC_OK;
}

int mpc_results_device(MpcEngine* e, const MpcResult** dev) {
  if (!e || !dev) return MPC_ERR_INVALID;
  if (!e->solved) return fail(e, MPC_ERR_STATE, "mpc_results_device before mpc_solve");
  *dev = e->d_results;
  return MPC_OK;
}

int mpc_get_solution(MpcEngine* e, int32_t idx, float* x) {
  if (!e || e->kind != 0 || !x) return MPC_ERR_INVALID;
  if (!e->solved) return fail(e, MPC_ERR_STATE, "mpc_get_solution before mpc_solve");
  if (idx < 0 || idx >= e->n) return fail(e, MPC_ERR_INVALID, "problem index out of range");
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  CUDA_TRY(e, cudaMemcpy(x, e->d_x + size_t(idx) * e->nvar(), e->nvar() * sizeof(float), cudaMemcpyDeviceToHost));
  return MPC_OK;
}

int mpc_compute_grf_batch(MpcEngine* e, const MpcStateIn* host_in, MpcResult* host_out, int32_t n) {
  int rc = mpc_load_states(e, host_in, n);
  if (rc) return rc;
  rc = mpc_build_qp_async(e);
  if (rc) return rc;
  rc = mpc_solve_async(e);
  if (rc) return rc;
  return mpc_get_results(e, host_out);
}

// ---- fleet: one box, several GPUs, results into one host array ------------------

}  // extern "C"

struct MpcFleet {
  std::vector<MpcEngine*> eng;
  std::vector<MpcStateIn*> pin_in;   // pinned staging per shard (pageable caller buffers only)
  std::vector<MpcResult*> pin_out;
  std::vector<int> pin_cap;
  std::string err;
};

namespace {

bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int fleet_fail(MpcFleet* f, int code, const std::string& msg) {
  if (f) f->err = msg;
  else g_create_error = msg;
  return code;
}

void shard_range(int n, int g, int i, int* b, int* e) {
  *b = int((int64_t(n) * i + g - 1) / g);
  *e = int((int64_t(n) * (i + 1) + g - 1) / g);
}

int fleet_run(MpcFleet* f, const MpcStateIn* host_in, MpcResult* host_out, int n, bool warm) {
  if (!f) return MPC_ERR_INVALID;
  if (n < 0 || (n > 0 && (!host_in || !host_out))) return fleet_fail(f, MPC_ERR_INVALID, "bad host buffers");
  const int G = int(f->eng.size());
  const bool in_pinned = n > 0 && host_ptr_is_pinned(host_in);
  const bool out_pinned = n > 0 && host_ptr_is_pinned(host_out);
  // enqueue every shard, then wait: the GPUs work and copy concurrently
  for (int i = 0; i < G; ++i) {
    int b, e2;
    shard_range(n, G, i, &b, &e2);
    const int m = e2 - b;
    MpcEngine* e = f->eng[i];
    if (m > f->pin_cap[i] && (!in_pinned || !out_pinned)) {
      if (cudaSetDevice(e->device) != cudaSuccess) return fleet_fail(f, MPC_ERR_CUDA, "cudaSetDevice");
      cudaStreamSynchronize(e->stream);
      cudaFreeHost(f->pin_in[i]);
      cudaFreeHost(f->pin_out[i]);
      f->pin_in[i] = nullptr;
      f->pin_out[i] = nullptr;
      f->pin_cap[i] = 0;
      if (cudaMallocHost(&f->pin_in[i], size_t(m) * sizeof(MpcStateIn)) != cudaSuccess ||
          cudaMallocHost(&f->pin_out[i], size_t(m) * sizeof(MpcResult)) != cudaSuccess)
        return fleet_fail(f, MPC_ERR_CUDA, "pinned staging allocation failed");
      f->pin_cap[i] = m;
    }
    const MpcStateIn* src = host_in + b;
    if (!in_pinned && m > 0) {
      std::memcpy(f->pin_in[i], src, size_t(m) * sizeof(MpcStateIn));
      src = f->pin_in[i];
    }
    int rc = mpc_load_states(e, src, m);
    if (rc == MPC_OK) rc = mpc_build_qp_async(e);
    if (rc == MPC_OK) rc = warm ? mpc_solve_warm_async(e) : mpc_solve_async(e);
    if (rc == MPC_OK && m > 0) {
      MpcResult* dst = out_pinned ? host_out + b : f->pin_out[i];
      if (cudaMemcpyAsync(dst, e->d_results, size_t(m) * sizeof(MpcResult), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess)
        rc = fleet_fail(f, MPC_ERR_CUDA, "result copy");
    }
    if (rc != MPC_OK) {
      if (f->err.empty() || rc != MPC_ERR_CUDA) f->err = "shard " + std::to_string(i) + ": " + e->err;
      return rc;
    }
  }
  for (int i = 0; i < G; ++i) {
    int b, e2;
    shard_range(n, G, i, &b, &e2);
    MpcEngine* e = f->eng[i];
    if (cudaSetDevice(e->device) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess)
      return fleet_fail(f, MPC_ERR_CUDA, "shard " + std::to_string(i) + ": synchronise failed");
    if (!out_pinned && e2 > b) std::memcpy(host_out + b, f->pin_out[i], size_t(e2 - b) * sizeof(MpcResult));
  }
  return MPC_OK;
}

}  // namespace

extern "C" {

int mpc_fleet_create(const MpcConfig* cfg, const int32_t* devices, int32_t ndev, MpcFleet** out) {
  if (!out) return fleet_fail(nullptr, MPC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!cfg || !devices || ndev <= 0 || ndev > 64) return fleet_fail(nullptr, MPC_ERR_INVALID, "bad device list");
  MpcFleet* f = new (std::nothrow) MpcFleet();
  if (!f) return fleet_fail(nullptr, MPC_ERR_INVALID, "out of host memory");
  for (int i = 0; i < ndev; ++i) {
    MpcEngine* e = nullptr;
    const int rc = mpc_engine_create(cfg, devices[i], &e);
    if (rc != MPC_OK) {
      const std::string msg = "device " + std::to_string(devices[i]) + ": " + g_create_error;
      mpc_fleet_destroy(f);
      return fleet_fail(nullptr, rc, msg);
    }
    f->eng.push_back(e);
    f->pin_in.push_back(nullptr);
    f->pin_out.push_back(nullptr);
    f->pin_cap.push_back(0);
  }
  *out = f;
  return MPC_OK;
}

void mpc_fleet_destroy(MpcFleet* f) {
  if (!f) return;
  for (size_t i = 0; i < f->eng.size(); ++i) {
    if (f->eng[i]) cudaSetDevice(f->eng[i]->device);
    if (f->eng[i] && f->eng[i]->stream) cudaStreamSynchronize(f->eng[i]->stream);
    cudaFreeHost(f->pin_in[i]);
    cudaFreeHost(f->pin_out[i]);
    mpc_engine_destroy(f->eng[i]);
  }
  delete f;
}

const char* mpc_fleet_last_error(const MpcFleet* f) { return f ? f->err.c_str() : g_create_error.c_str(); }

int32_t mpc_fleet_size(const MpcFleet* f) { return f ? int32_t(f->eng.size()) : 0; }

int mpc_fleet_shard_range(int32_t n, int32_t ndev, int32_t i, int32_t* begin, int32_t* end) {
  if (n < 0 || ndev <= 0 || i < 0 || i >= ndev || !begin || !end) return MPC_ERR_INVALID;
  int b, e;
  shard_range(n, ndev, i, &b, &e);
  *begin = b;
  *end = e;
  return MPC_OK;
}

int mpc_fleet_compute_grf_batch(MpcFleet* f, const MpcStateIn* host_in, MpcResult* host_out, int32_t n) {
  return fleet_run(f, host_in, host_out, n, false);
}

int mpc_fleet_stream_step(MpcFleet* f, const MpcStateIn* host_in, MpcResult* host_out, int32_t n) {
  return fleet_run(f, host_in, host_out, n, true);
}

int mpc_fleet_stream_reset(MpcFleet* f) {
  if (!f) return MPC_ERR_INVALID;
  for (MpcEngine* e : f->eng) {
    const int rc = mpc_stream_reset(e);
    if (rc) return fleet_fail(f, rc, e->err);
  }
  return MPC_OK;
}

int64_t mpc_fleet_kernel_launches(const MpcFleet* f) {
  int64_t s = 0;
  if (f) for (const MpcEngine* e : f->eng) s += e->launches;
  return s;
}

// ---- ConvexMpc surface, one problem -----------------------------------------

int mpc_qp_mats_from_model(MpcEngine* e, const double* A_mat_d, const double* B_mat_d_list,
                           const double* mpc_states, const double* mpc_states_d,
                           const int32_t* contacts, double* hessian, double* gradient, double* lb,
                           double* ub) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!A_mat_d || !B_mat_d_list || !mpc_states || !mpc_states_d || !contacts)
    return fail(e, MPC_ERR_INVALID, "NULL model input");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const int Hh = e->H, nv = e->nvar(), nc = e->ncon(), ns = 13 * e->H;
  const size_t rs = (Hh == kH) ? size_t(kNP) : size_t(nv);  // row stride of P on the device
  const size_t nd = 169 + size_t(Hh) * 156 + 13 + ns;
  const size_t npq = e->p_stride() + nv;  // P | q, f64
  double* d_model = nullptr;
  int* d_contacts = nullptr;
  double* d_pq = nullptr;
  float* d_lu = nullptr;
  CUDA_TRY(e, cudaMalloc(&d_model, nd * sizeof(double)));
  cudaError_t crc = cudaMalloc(&d_contacts, 4 * sizeof(int));
  if (crc == cudaSuccess) crc = cudaMalloc(&d_pq, npq * sizeof(double));
  if (crc == cudaSuccess) crc = cudaMalloc(&d_lu, 2 * nc * sizeof(float));
  std::vector<double> hm(nd);
  std::memcpy(hm.data(), A_mat_d, 169 * sizeof(double));
  std::memcpy(hm.data() + 169, B_mat_d_list, size_t(Hh) * 156 * sizeof(double));
  std::memcpy(hm.data() + 169 + Hh * 156, mpc_states, 13 * sizeof(double));
  std::memcpy(hm.data() + 169 + Hh * 156 + 13, mpc_states_d, ns * sizeof(double));
  if (crc == cudaSuccess)
    crc = cudaMemcpyAsync(d_model, hm.data(), nd * sizeof(double), cudaMemcpyHostToDevice, e->stream);
  if (crc == cudaSuccess)
    crc = cudaMemcpyAsync(d_contacts, contacts, 4 * sizeof(int), cudaMemcpyHostToDevice, e->stream);
  int rc = MPC_OK;
  std::vector<double> hpq(npq);
  std::vector<float> hlu(2 * nc);
  if (crc == cudaSuccess) {
    ModelIn m{};
    m.A_d = d_model;
    m.B_d_list = d_model + 169;
    m.x0 = d_model + 169 + Hh * 156;
    m.x_ref = d_model + 169 + Hh * 156 + 13;
    m.contacts = d_contacts;
    rc = launch_build(e, nullptr, m, 1, d_pq, d_pq + e->p_stride(), d_lu, d_lu + nc);
    if (rc == MPC_OK) {
      crc = cudaMemcpyAsync(hpq.data(), d_pq, npq * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
      if (crc == cudaSuccess)
        crc = cudaMemcpyAsync(hlu.data(), d_lu, 2 * nc * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
      if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
    }
  }
  cudaFree(d_model);
  cudaFree(d_contacts);
  cudaFree(d_pq);
  cudaFree(d_lu);
  if (rc) return rc;
  if (crc != cudaSuccess) return fail(e, MPC_ERR_CUDA, cudaGetErrorString(crc));
  if (hessian)
    for (int r = 0; r < nv; ++r)
      for (int c = 0; c < nv; ++c) hessian[size_t(r) * nv + c] = hpq[size_t(r) * rs + c];
  if (gradient) for (int i = 0; i < nv; ++i) gradient[i] = hpq[e->p_stride() + i];
  // the device stores bounds in fp32; hand back exactly +-OsqpEigen::INFTY like the reference
  auto snap = [](float v) -> double { return v >= 1e29f ? MPC_INFTY : (v <= -1e29f ? -MPC_INFTY : (double)v); };
  if (lb) for (int i = 0; i < nc; ++i) lb[i] = snap(hlu[i]);
  if (ub) for (int i = 0; i < nc; ++i) ub[i] = snap(hlu[nc + i]);
  return MPC_OK;
}

int mpc_solve_qp(MpcEngine* e, const double* hessian, const double* gradient, const double* lb,
                 const double* ub, double* solution, int32_t* status, int32_t* iters) {
  if (!e || e->kind != 0) return MPC_ERR_INVALID;
  if (!hessian || !gradient || !lb || !ub || !solution) return fail(e, MPC_ERR_INVALID, "NULL QP input");
  // The device solvers do not evaluate OSQP's infeasibility certificates (the QPs of this path are feasible
  // and strictly convex by construction); a caller-supplied QP with crossed bounds is refused here like
  // osqp_setup / osqp_update_bounds refuse it, instead of burning max_iter iterations on it.
  for (int i = 0; i < e->ncon(); ++i)
    if (!(lb[i] <= ub[i])) return fail(e, MPC_ERR_INVALID, "mpc_solve_qp: lb[" + std::to_string(i) + "] > ub (or NaN)");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const int Hh = e->H, nv = e->nvar(), nc = e->ncon();
  const size_t rs = (Hh == kH) ? size_t(kNP) : size_t(nv);
  const size_t npq = e->p_stride() + nv;
  std::vector<double> hpq(npq, 0.0);
  for (int r = 0; r < nv; ++r)
    for (int c = 0; c < nv; ++c) hpq[size_t(r) * rs + c] = hessian[size_t(r) * nv + c];
  for (int i = 0; i < nv; ++i) hpq[e->p_stride() + i] = gradient[i];
  std::vector<float> hlu(2 * nc);
  for (int i = 0; i < nc; ++i) { hlu[i] = (float)lb[i]; hlu[nc + i] = (float)ub[i]; }
  double* d_pq = nullptr;
  float* d_lu = nullptr;
  float* d_xs = nullptr;
  MpcResult* d_res = nullptr;
  CUDA_TRY(e, cudaMalloc(&d_pq, npq * sizeof(double)));
  cudaError_t crc = cudaMalloc(&d_lu, 2 * nc * sizeof(float));
  if (crc == cudaSuccess) crc = cudaMalloc(&d_xs, nv * sizeof(float));
  if (crc == cudaSuccess) crc = cudaMalloc(&d_res, sizeof(MpcResult));
  if (crc == cudaSuccess)
    crc = cudaMemcpyAsync(d_pq, hpq.data(), npq * sizeof(double), cudaMemcpyHostToDevice, e->stream);
  if (crc == cudaSuccess)
    crc = cudaMemcpyAsync(d_lu, hlu.data(), 2 * nc * sizeof(float), cudaMemcpyHostToDevice, e->stream);
  int rc = MPC_OK;
  std::vector<float> hx(nv);
  MpcResult hr{};
  if (crc == cudaSuccess) {
    rc = launch_solve(e, d_pq, d_pq + e->p_stride(), d_lu, d_lu + nc, nullptr, d_res, d_xs, 1);
    if (rc == MPC_OK) {
      crc = cudaMemcpyAsync(hx.data(), d_xs, nv * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
      if (crc == cudaSuccess)
        crc = cudaMemcpyAsync(&hr, d_res, sizeof(MpcResult), cudaMemcpyDeviceToHost, e->stream);
      if (crc == cudaSuccess) crc = cudaStreamSynchronize(e->stream);
    }
  }
  cudaFree(d_pq);
  cudaFree(d_lu);
  cudaFree(d_xs);
  cudaFree(d_res);
  if (rc) return rc;
  if (crc != cudaSuccess) return fail(e, MPC_ERR_CUDA, cudaGetErrorString(crc));
  for (int i = 0; i < nv; ++i) solution[i] = hx[i];
  if (status) *status = hr.status;
  if (iters) *iters = hr.iters;
  return MPC_OK;
}

// ---- stance-balance QP --------------------------------------------------------

int balance_load_states(MpcEngine* e, const BalanceStateIn* host, int32_t n) {
  if (!e || e->kind != 1) return MPC_ERR_INVALID;
  if (n < 0 || (n > 0 && !host)) return fail(e, MPC_ERR_INVALID, "bad state buffer");
  CUDA_TRY(e, cudaSetDevice(e->device));
  int rc = reserve(e, n);
  if (rc) return rc;
  if (n > 0)
    CUDA_TRY(e, cudaMemcpyAsync(e->d_bstates, host, size_t(n) * sizeof(BalanceStateIn),
                                cudaMemcpyHostToDevice, e->stream));
  e->n = n;
  e->built = true;  // build and solve are one fused kernel for the 12-variable QP
  e->solved = false;
  e->torque_on = false;
  return MPC_OK;
}

int balance_solve(MpcEngine* e) {
  if (!e || e->kind != 1) return MPC_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  if (e->n > 0) {
    CUDA_TRY(e, cudaMemsetAsync(e->d_counter, 0, sizeof(int), e->stream));
    static const bool warp_kernel = [] { const char* v = std::getenv("MPC_BALANCE_KERNEL"); return v && v[0] == 'w'; }();
    if (warp_kernel) {
      // one warp per problem (round 1's layout, kept as the second implementation: MPC_BALANCE_KERNEL=warp)
      const int warps_per_cta = kBalanceThreads / 32;
      const int need = (e->n + warps_per_cta - 1) / warps_per_cta, full = e->num_sms * kBalanceCtasPerSm;
      const int grid = need < full ? need : full;
      balance_qp_kernel<<<grid, kBalanceThreads, 0, e->stream>>>(e->d_bstates, e->n, e->d_counter, e->d_Pb, e->d_qb,
                                                                 e->d_l, e->d_u, e->d_results, e->bal);
    } else {
      // four lanes per problem, 32 problems per CTA
      const int per_cta = kBalLegThreads / 4;
      const int need = (e->n + per_cta - 1) / per_cta, full = e->num_sms * kBalLegCtasPerSm;
      const int grid = need < full ? need : full;
      balance_qp_leg_kernel<<<grid, kBalLegThreads, per_cta * kBalLegPStride * sizeof(double), e->stream>>>(
          e->d_bstates, e->n, e->d_counter, e->d_Pb, e->d_qb, e->d_l, e->d_u, e->d_results, e->bal);
    }
    ++e->launches;
    CUDA_TRY(e, cudaGetLastError());
    if (e->torque_on) {
      torque_map_kernel<<<(4 * e->n + 127) / 128, 128, 0, e->stream>>>(
          e->d_results, reinterpret_cast<const float*>(e->d_bstates), int(sizeof(BalanceStateIn) / 4), 54, e->d_tin,
          e->d_tout, e->n);
      ++e->launches;
      CUDA_TRY(e, cudaGetLastError());
    }
  }
  e->solved = true;
  return MPC_OK;
}

int balance_get_qp(MpcEngine* e, int32_t idx, float* P, float* q, float* l, float* u) {
  if (!e || e->kind != 1) return MPC_ERR_INVALID;
  if (!e->solved) return fail(e, MPC_ERR_STATE, "balance_get_qp before balance_solve");
  if (idx < 0 || idx >= e->n) return fail(e, MPC_ERR_INVALID, "problem index out of range");
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (P) CUDA_TRY(e, cudaMemcpy(P, e->d_Pb + size_t(idx) * 144, 144 * sizeof(float), cudaMemcpyDeviceToHost));
  if (q) CUDA_TRY(e, cudaMemcpy(q, e->d_qb + size_t(idx) * 12, 12 * sizeof(float), cudaMemcpyDeviceToHost));
  if (l) CUDA_TRY(e, cudaMemcpy(l, e->d_l + size_t(idx) * 20, 20 * sizeof(float), cudaMemcpyDeviceToHost));
  if (u) CUDA_TRY(e, cudaMemcpy(u, e->d_u + size_t(idx) * 20, 20 * sizeof(float), cudaMemcpyDeviceToHost));
  return MPC_OK;
}

int balance_qp_solve(MpcEngine* e, const BalanceStateIn* host_in, MpcResult* host_out, int32_t n) {
  int rc = balance_load_states(e, host_in, n);
  if (rc) return rc;
  rc = balance_solve(e);
  if (rc) return rc;
  return mpc_get_results(e, host_out);
}

}  // extern "C"
