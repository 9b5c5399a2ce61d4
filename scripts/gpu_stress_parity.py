"""Wider parity sweep of the two kernels added late in round 2 (run under gpurun): wrench_riccati_kernel<30> on fresh
seeds and both weight sets, balance_qp_leg_kernel on 65536 states against the oracle."""
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


for name in ("gazebo", "hardware"):
    for seed in (2001, 2002):
        cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
        cfg.horizon = 30
        n = 1024
        st = pkg.generate_states(seed, 0, n)
        e = pkg.MpcEngine(cfg, 0)
        r = e.compute_grf_batch(st)
        ref = ob.mpc_compute_grf(cfg, st)
        print(f"[H=30 {name} seed {seed}] status ok {(r['status'] == ref['status']).all()} same iters {(r['iters'] == ref['iters']).mean():.4f} "
              f"same rho {(r['rho_updates'] == ref['rho_updates']).mean():.4f} max GRF rel {rel(r['grf'], ref['grf']).max():.2e}", flush=True)
        e.close()
bcfg = pkg.balance_config_default()
be = pkg.MpcEngine(bcfg, 0, balance=True)
for seed in (3001, 3002):
    n = 65536
    st = pkg.generate_balance_states(seed, 0, n)
    r = be.compute_grf_batch(st)
    ref = ob.balance_compute_grf(bcfg, st)
    same = r["iters"] == ref["iters"]
    print(f"[balance seed {seed}] status ok {(r['status'] == ref['status']).all()} same iters {same.mean():.5f} flipped {int((~same).sum())} "
          f"max GRF rel (same) {rel(r['grf'], ref['grf'])[same].max():.2e}", flush=True)
be.close()
print("done")
