"""Small end-to-end run for compute-sanitizer: every kernel of the library once (cold, warm,
H = 30, stance QP, state preparation, torque map, horizon extensions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = pkg.config_default()
cfg.exact_discretization = cfg.foot_drift = cfg.gait_aware = 1
e = pkg.MpcEngine(cfg, 0)
for t in range(3):
    e.prepare_states(pkg.generate_sensors(1002, 0, n, t))
    e.set_gait_inputs(pkg.generate_gait_inputs(1002, 0, n, t))
    e.build_qp(sync=False)
    e.solve_warm(sync=False)
    r = e.get_results(); tq = e.get_torques()
print("warm pipeline", r["iters"].mean(), (r["status"] == 1).all(), int((tq["nan_mask"] != 0).sum()))
e.close()
e = pkg.MpcEngine(pkg.config_default(), 0)
r = e.compute_grf_batch(pkg.generate_states(1002, 0, n))
print("cold", r["iters"].mean(), (r["status"] == 1).all())
e.close()
c30 = pkg.config_default(); c30.horizon = 30
e = pkg.MpcEngine(c30, 0)
r = e.compute_grf_batch(pkg.generate_states(1004, 0, 4))
print("h30", r["iters"].mean(), (r["status"] == 1).all())
e.close()
e = pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True)
r = e.compute_grf_batch(pkg.generate_balance_states(1005, 0, n))
print("balance", r["iters"].mean())
e.close()
