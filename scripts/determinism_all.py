"""Run-to-run determinism of every solver path: warm stream, pipeline, H = 30, stance QP."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, go1_qp_mpc_controller_b200 as pkg
R = int(sys.argv[1]) if len(sys.argv) > 1 else 12
def same(a, b): return all(np.array_equal(a[k], b[k]) for k in ("iters", "grf", "status"))
# warm stream: R fresh engines, 6 ticks each
n, T = 1024, 6
st = [pkg.generate_stream_states(1002, 0, n, 44 + t) for t in range(T)]
ref = None; bad = 0
for r in range(R):
    e = pkg.MpcEngine(pkg.config_default(), 0)
    out = [e.stream_step(s).copy() for s in st]
    if ref is None: ref = out
    elif not all(same(a, b) for a, b in zip(out, ref)): bad += 1
    e.close()
print("warm stream runs differing:", bad, "of", R - 1)
# pipeline with all extensions
cfg = pkg.config_default(); cfg.exact_discretization = cfg.foot_drift = cfg.gait_aware = 1
ref = None; bad = 0
for r in range(R):
    e = pkg.MpcEngine(cfg, 0); outs = []
    for t in range(4):
        e.prepare_states(pkg.generate_sensors(1002, 0, n, t)); e.set_gait_inputs(pkg.generate_gait_inputs(1002, 0, n, t))
        e.build_qp(sync=False); e.solve_warm(sync=False); outs.append((e.get_results().copy(), e.get_torques().copy()))
    if ref is None: ref = outs
    elif not all(same(a[0], b[0]) and np.array_equal(a[1]["joint_torques"], b[1]["joint_torques"]) for a, b in zip(outs, ref)): bad += 1
    e.close()
print("pipeline runs differing:", bad, "of", R - 1)
c30 = pkg.config_default(); c30.horizon = 30
s30 = pkg.generate_states(1004, 0, 160); ref = None; bad = 0
for r in range(4):
    e = pkg.MpcEngine(c30, 0); out = e.compute_grf_batch(s30).copy()
    if ref is None: ref = out
    elif not same(out, ref): bad += 1
    e.close()
print("h30 runs differing:", bad, "of 3")
sb = pkg.generate_balance_states(1005, 0, 8192); ref = None; bad = 0
for r in range(R):
    e = pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True); out = e.compute_grf_batch(sb).copy()
    if ref is None: ref = out
    elif not same(out, ref): bad += 1
    e.close()
print("balance runs differing:", bad, "of", R - 1)
