"""Experiment (north_star's "Hessian on tensor cores"): the condensed Hessian
P = B_qp' Q B_qp + R formed ON THE GPU'S TENSOR CORES in TF32 (1x and error-compensated 3x, cuBLAS batched
GEMM, allow_tf32) and in plain FP32, then every QP solved by the f64 dense solver (mpc_solve_qp) and the
first-step forces compared with the solve on the exact f64 Hessian.  Measures the failure rate of the 1e-3
GRF gate that a reduced-precision Hessian causes.  Library GEMMs are fine here: this is an experiment about
arithmetic, not the product path (which, since round 2, never forms a Hessian at all)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = pkg.config_default()
states = pkg.generate_states(1002, 0, N)
t0 = time.time()
Bq = np.empty((N, 130, 120))
qv = np.empty((N, 120))
lo = np.empty((N, 200))
hi = np.empty((N, 200))
for i in range(N):
    it = ob.mpc_build_intermediates(cfg, states[i])
    Bq[i] = it["B_qp"]
    _, q, l, u = ob.mpc_build_qp(cfg, states[i])
    qv[i], lo[i], hi[i] = q, l, u
print(f"oracle intermediates for {N} states: {time.time() - t0:.1f} s", flush=True)
Q = torch.tensor(np.tile(2.0 * np.array(cfg.q_weights[:]), 10), dtype=torch.float64, device="cuda")
R = torch.tensor(np.tile(2.0 * np.array(cfg.r_weights[:]), 10), dtype=torch.float64, device="cuda")
B = torch.tensor(Bq, device="cuda")                       # (N, 130, 120) f64
QB = Q[None, :, None] * B


def tf32_round(x32):
    """zero the 13 low mantissa bits (round to nearest even on the kept 10 bits)"""
    i = x32.view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) / reps


variants = {}
torch.backends.cuda.matmul.allow_tf32 = False
variants["f64 (exact)"] = timed(lambda: torch.bmm(B.transpose(1, 2), QB) + torch.diag(R)[None])
A32, C32 = B.transpose(1, 2).contiguous().float(), QB.float()
variants["f32 FFMA"] = timed(lambda: torch.bmm(A32, C32).double() + torch.diag(R)[None])
torch.backends.cuda.matmul.allow_tf32 = True
variants["TF32 x1 (tensor cores)"] = timed(lambda: torch.bmm(A32, C32).double() + torch.diag(R)[None])
Ah, Ch = tf32_round(A32), tf32_round(C32)
Al, Cl = A32 - Ah, C32 - Ch
variants["TF32 x3 (tensor cores, hi/lo split)"] = timed(
    lambda: (torch.bmm(Ah, Ch) + (torch.bmm(Ah, Cl) + torch.bmm(Al, Ch))).double() + torch.diag(R)[None])
torch.backends.cuda.matmul.allow_tf32 = False

eng = pkg.MpcEngine(cfg, 0)
Pex = variants["f64 (exact)"][0].cpu().numpy()
sol_ex = np.empty((N, 12))
it_ex = np.empty(N, dtype=int)
for i in range(N):
    x, st, it = eng.solve_qp(Pex[i], qv[i], lo[i], hi[i])
    sol_ex[i], it_ex[i] = x[:12], it
print("variant | GEMM ms per %d problems | max|dP|/max|P| | same iteration count | GRF rel err max / median | states past the 1e-3 gate" % N)
for name, (P, sec) in variants.items():
    Pn = P.cpu().numpy()
    dP = np.abs(Pn - Pex).max(axis=(1, 2)) / np.abs(Pex).max(axis=(1, 2))
    if name.startswith("f64"):
        print(f"{name:38s} | {sec * 1e3:8.3f} | {dP.max():.1e} | reference")
        continue
    rel = np.empty(N)
    same = 0
    for i in range(N):
        Ps = 0.5 * (Pn[i] + Pn[i].T)
        x, st, it = eng.solve_qp(Ps, qv[i], lo[i], hi[i])
        rel[i] = np.linalg.norm(x[:12] - sol_ex[i]) / max(np.linalg.norm(sol_ex[i]), 1.0)
        same += it == it_ex[i]
    print(f"{name:38s} | {sec * 1e3:8.3f} | {dP.max():.1e} | {same / N:.4f} | {rel.max():.2e} / {np.median(rel):.2e} | "
          f"{(rel > 1e-3).sum()} of {N} ({100 * (rel > 1e-3).mean():.2f} %)", flush=True)
eng.close()
