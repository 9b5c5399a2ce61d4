"""Small fixed workload for ncu: build + solve of 296 states (2 CTAs per SM-slot)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
eng = pkg.MpcEngine(pkg.config_default(), 0)
st = pkg.generate_states(1002, 0, n)
for _ in range(2):
    res = eng.compute_grf_batch(st)
print("ok", res["iters"].mean(), eng.kernel_launches())
