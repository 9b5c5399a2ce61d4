"""Stance-balance QP engine: one timed batch (the ncu target) + throughput at the per-GPU share of config 5."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
be = pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True)
st = pkg.generate_balance_states(1005, 0, n)
r = be.compute_grf_batch(st)
t0 = time.perf_counter()
for _ in range(3):
    r = be.compute_grf_batch(st)
dt = (time.perf_counter() - t0) / 3
print(f"balance n={n}: {dt * 1e3:.2f} ms -> {n / dt / 1e6:.2f} M/s, iters mean {r['iters'].mean():.1f} max {r['iters'].max()}, solved {(r['status'] == 1).mean():.4f}")
be.close()
