"""The widened path, sensors in -> torques out, per control tick:
state preparation (orientation, leg kinematics, EKF, terrain pitch) -> QP build -> warm-started
solve -> GRF + joint torques.  n robots, one GPU, host buffers in and out every tick."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 24
eng = pkg.MpcEngine(pkg.config_default(), 0)
pcfg = pkg.prep_config_default()
import torch  # pinned host buffers only (plumbing), as in bench.py's e2e leg
def pinned(dtype, count):
    t = torch.empty(count * dtype.itemsize, dtype=torch.uint8).pin_memory()
    return t, t.numpy().view(dtype)
_keep = []
sens = []
for t in range(ticks):
    buf, arr = pinned(pkg.abi.SENSOR_DTYPE, n)
    arr[:] = pkg.generate_sensors(1002, 0, n, 40 + t)
    _keep.append(buf); sens.append(arr)
_rb, res_buf = pinned(pkg.abi.RESULT_DTYPE, n)
eng.prepare_states(sens[0][:296], pcfg); eng.build_qp(); eng.solve_warm(); eng.get_results()
eng.prepare_reset(); eng.stream_reset()
rows = []
for t in range(ticks):
    t0 = time.perf_counter()
    eng.prepare_states(sens[t], pcfg)
    t1 = time.perf_counter()
    eng.build_qp(sync=False)
    eng.solve_warm(sync=False)
    res = eng.get_results(res_buf)
    tq = eng.get_torques()
    t2 = time.perf_counter()
    eng.synchronize()
    rows.append({"tick": t, "ms": 1e3 * (t2 - t0), "enqueue_prep_ms": 1e3 * (t1 - t0), "mean_iters": float(res["iters"].mean()),
                 "solved": float((res["status"] == 1).mean()), "nan_torques": int((tq["nan_mask"] != 0).sum())})
# the preparation kernel alone
eng.prepare_states(sens[-1], pcfg); eng.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    eng.prepare_states(sens[-1], pcfg)
eng.synchronize()
prep_ms = 1e3 * (time.perf_counter() - t0) / 10
warm = rows[2:]
med = float(np.median([r["ms"] for r in warm]))
print(json.dumps({"what": "sensors -> state prep -> build -> warm solve -> GRF + torques, one GPU, host to host per tick",
                  "robots": n, "ticks": ticks, "tick_ms_median": med, "robot_ticks_per_s": n / med * 1e3,
                  "state_prep_ms_incl_h2d": prep_ms, "warm_mean_iters": float(np.mean([r["mean_iters"] for r in warm])),
                  "all_solved": all(r["solved"] == 1.0 for r in rows), "per_tick": rows}))
