"""Throughput of the default H = 10 engine alone (cold batches of 4096 and 16384 states, host to host)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

e = pkg.MpcEngine(pkg.config_default(), 0)
for n in (4096, 16384):
    batches = [pkg.generate_states(1234, b * n, n) for b in range(6)]
    for b in batches[:2]:
        r = e.compute_grf_batch(b)
    t = time.perf_counter()
    for b in batches[2:]:
        r = e.compute_grf_batch(b)
    dt = time.perf_counter() - t
    print(f"n={n}: {4 * n / dt / 1e3:.1f} k solves/s, solved {int((r['status'] == 1).sum())}, mean iters {r['iters'].mean():.1f}")
e.close()
