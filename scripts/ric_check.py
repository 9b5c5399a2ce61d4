"""Parity and timing of the Riccati-structured solver against the oracle (and the dense kernels)."""
import sys, time; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, go1_qp_mpc_controller_b200 as pkg, oracle_binding as ob
for H, n, seed in ((10, 4096, 1002), (30, 2048, 1004)):
    cfg = pkg.config_default(); cfg.horizon = H; cfg.structured_solver = 1
    e = pkg.MpcEngine(cfg, 0)
    st = pkg.generate_states(seed, 0, n)
    res = e.compute_grf_batch(st)
    t0 = time.perf_counter(); res = e.compute_grf_batch(st); dt = time.perf_counter() - t0
    k = min(n, 256 if H == 10 else 64)
    ref = ob.mpc_compute_grf(cfg, st[:k])
    den = np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
    rel = np.linalg.norm(res["grf"][:k].astype(np.float64) - ref["grf"], axis=1) / den
    print(f"H={H}: {n/dt:.0f} solves/s, status ok {(res['status']==1).mean():.3f}, same iters {(res['iters'][:k]==ref['iters']).mean():.4f}, "
          f"max rel grf err {rel.max():.2e}, mean iters {res['iters'].mean():.1f} vs {ref['iters'].mean():.1f}")
    e.close()
