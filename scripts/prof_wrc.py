"""Batches of H = 30 states through the long-horizon wrench-space engine: the ncu / WRC_PROF target."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = pkg.config_default()
cfg.horizon = 30
cfg.structured_solver = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = pkg.MpcEngine(cfg, 0)
for b in range(3):
    r = e.compute_grf_batch(pkg.generate_states(1004, b * n, n))
print("ok", int((r["status"] == 1).sum()), float(r["iters"].mean()))
e.close()
