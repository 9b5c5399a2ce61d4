"""GPU check of the balance QP kernels: parity of the four-lanes-per-problem kernel against the oracle and against the
one-warp-per-problem kernel (MPC_BALANCE_KERNEL=warp in a second process), timing.  Run under gpurun."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

which = os.environ.get("MPC_BALANCE_KERNEL", "leg")
bcfg = pkg.balance_config_default()
be = pkg.MpcEngine(bcfg, 0, balance=True)
n = 8192
st = pkg.generate_balance_states(1005, 0, n)
res = be.compute_grf_batch(st).copy()
if "--dump" in sys.argv:
    np.save(os.path.join(ROOT, "gpurun_out", f"balance_{which}.npy"), res)
if "--oracle" in sys.argv:
    import oracle_binding as ob
    ref = ob.balance_compute_grf(bcfg, st)
    same = res["iters"] == ref["iters"]
    rel = np.linalg.norm(res["grf"].astype(np.float64) - ref["grf"], axis=1) / np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
    print(f"[{which}] status equal {np.array_equal(res['status'], ref['status'])}, same iters {same.mean():.4f}, "
          f"same rho updates {(res['rho_updates'] == ref['rho_updates']).mean():.4f}, max GRF rel (same iters) {rel[same].max():.2e}, "
          f"overall {rel.max():.2e}", flush=True)
    P, q, l, u = be.get_qp(3)
    print(f"[{which}] get_qp finite {np.isfinite(P).all()} l {l[:6]} u {u[:6]}")
for nb in (32768, 125000):
    stb = pkg.generate_balance_states(1005, 0, nb)
    out = np.zeros(nb, dtype=pkg.abi.RESULT_DTYPE)
    be.compute_grf_batch(stb, out)
    t0 = time.perf_counter()
    for _ in range(3):
        be.compute_grf_batch(stb, out)
    dt = (time.perf_counter() - t0) / 3
    print(f"[{which}] n={nb}: {dt * 1e3:.2f} ms host to host (pageable) -> {nb / dt / 1e6:.2f} M/s, mean iters {out['iters'].mean():.1f}, "
          f"max {out['iters'].max()}", flush=True)
be.close()
