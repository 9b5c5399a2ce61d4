"""Three cold batches of 4096 states through the default (wrench-space) engine: the ncu target."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = pkg.MpcEngine(pkg.config_default(), 0)
for b in range(3):
    r = e.compute_grf_batch(pkg.generate_states(1002, b * n, n))
print("ok", int((r["status"] == 1).sum()), float(r["iters"].mean()))
e.close()
