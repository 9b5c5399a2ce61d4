"""OSQP settings sweep of the H = 30 wrench engine against the oracle (run under gpurun)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402


def rel(a, b):
    return np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1.0)


cases = {
    "scaling 0": dict(scaling=0),
    "adaptive_rho off": dict(adaptive_rho=0),
    "eps 1e-3": dict(eps_abs=1e-3, eps_rel=1e-3),
    "max_iter 60": dict(max_iter=60),
    "max_iter 110": dict(max_iter=110),
    "alpha 1.0": dict(alpha=1.0),
    "check 10, adapt 20": dict(check_termination=10, adaptive_rho_interval=20),
    "rho 1.0": dict(rho=1.0),
    "scaling 3": dict(scaling=3),
}
n = 96
st = pkg.generate_states(1004, 0, n)
for H in (30, 10):
    for name, kw in cases.items():
        cfg = pkg.config_hardware()
        cfg.horizon = H
        for k, v in kw.items():
            setattr(cfg.osqp, k, v)
        try:
            e = pkg.MpcEngine(cfg, 0)
        except Exception as ex:  # noqa: BLE001
            print(f"[H={H} {name}] engine refused: {ex}")
            continue
        r = e.compute_grf_batch(st)
        ref = ob.mpc_compute_grf(cfg, st)
        same = r["iters"] == ref["iters"]
        print(f"[H={H} {name}] status equal {np.array_equal(r['status'], ref['status'])} same iters {same.mean():.3f} "
              f"same rho {(r['rho_updates'] == ref['rho_updates']).mean():.3f} max GRF rel {rel(r['grf'], ref['grf']).max():.2e} "
              f"status set {sorted(set(r['status'].tolist()))} mean iters {r['iters'].mean():.1f}", flush=True)
        e.close()
print("done")
