"""Wider parity sweep of the H = 10 default kernel (wrench_tile_kernel) against the oracle: fresh seeds, both weight
sets, 4096 states each, and the warm stream over 12 ticks.  Run under gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402

ob.use_native()


def rel(a, b):
    return np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1.0)


for name in ("gazebo", "hardware"):
    for seed in (4001, 4002):
        cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
        n = 4096
        st = pkg.generate_states(seed, 0, n)
        e = pkg.MpcEngine(cfg, 0)
        r = e.compute_grf_batch(st)
        ref = ob.mpc_compute_grf(cfg, st)
        four = st["contacts"].sum(axis=1) == 4
        rr = rel(r["grf"], ref["grf"])
        print(f"[H=10 {name} seed {seed}] status ok {(r['status'] == ref['status']).all()} same iters {(r['iters'] == ref['iters']).mean():.5f} "
              f"same rho {(r['rho_updates'] == ref['rho_updates']).mean():.5f} max GRF rel {rr.max():.2e} (four-stance {rr[four].max():.2e}, "
              f"others {rr[~four].max():.2e})", flush=True)
        e.close()
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    T, n = 12, 512
    st = np.stack([pkg.generate_stream_states(4003, 0, n, 30 + t) for t in range(T)])
    ref = ob.mpc_stream(cfg, st)
    e = pkg.MpcEngine(cfg, 0)
    same, worst = [], 0.0
    for t in range(T):
        r = e.stream_step(st[t])
        same.append(float((r["iters"] == ref["iters"][t]).mean()))
        worst = max(worst, float(rel(r["grf"], ref["grf"][t]).max()))
    print(f"[H=10 {name} warm stream 12 ticks x {n}] same iters per tick min {min(same):.5f}, max GRF rel {worst:.2e}", flush=True)
    e.close()
print("done")
