"""One batch each through the kernels that had no ncu record: stance-balance QP, state preparation,
long-horizon build + Riccati solve.  The ncu target of profiles/r02_misc_ncu_summary.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

be = pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True)
for _ in range(2):
    r = be.compute_grf_batch(pkg.generate_balance_states(1005, 0, 32768))
print("balance", float(r["iters"].mean()))
be.close()
e = pkg.MpcEngine(pkg.config_default(), 0)
cfg = pkg.prep_config_default()
for t in range(2):
    e.prepare_states(pkg.generate_sensors(1007, 0, 4096, 3 * t), cfg)
    e.build_qp()
    e.solve_warm()
print("prep + warm", float(e.get_results()["iters"].mean()))
e.close()
c30 = pkg.config_default()
c30.horizon = 30
e30 = pkg.MpcEngine(c30, 0)
for _ in range(2):
    r = e30.compute_grf_batch(pkg.generate_states(1004, 0, 296))
print("h30", float(r["iters"].mean()))
e30.close()
