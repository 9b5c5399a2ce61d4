import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg
H = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 592
cfg = pkg.config_default(); cfg.horizon = H; cfg.structured_solver = 1
eng = pkg.MpcEngine(cfg, 0)
st = pkg.generate_states(1002 if H == 10 else 1004, 0, n)
for _ in range(2):
    res = eng.compute_grf_batch(st)
print("ok", res["iters"].mean(), eng.kernel_launches())
