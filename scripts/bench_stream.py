"""Warm-started streaming: n robots, consecutive control ticks, one
persistent solver per robot slot (A1RobotControl.cpp:522-538).  Host to host per tick."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 24
cfg = pkg.config_default()
eng = pkg.MpcEngine(cfg, 0)
import torch  # pinned host buffers only (plumbing), as in bench.py's e2e leg
def pinned(dtype, count):
    t = torch.empty(count * dtype.itemsize, dtype=torch.uint8).pin_memory()
    return t, t.numpy().view(dtype)
_keep = []
states = []
for t in range(ticks):
    buf, arr = pinned(pkg.abi.STATE_DTYPE, n)
    arr[:] = pkg.generate_stream_states(1002, 0, n, 40 + t)
    _keep.append(buf); states.append(arr)
eng.compute_grf_batch(states[0][:296])
eng.stream_reset()
_ob, out = pinned(pkg.abi.RESULT_DTYPE, n)
rows = []
for t in range(ticks):
    t0 = time.perf_counter(); eng.stream_step(states[t], out); dt = time.perf_counter() - t0
    rows.append({"tick": t, "ms": 1e3 * dt, "mean_iters": float(out["iters"].mean()),
                 "mean_factorisations": 1 + float(out["rho_updates"].mean()), "solved": float((out["status"] == 1).mean())})
t0 = time.perf_counter(); cold = eng.compute_grf_batch(states[-1]); cold_dt = time.perf_counter() - t0
import oracle_binding as ob
k, kt = 256, 6
t0 = time.perf_counter(); ref = ob.mpc_stream(cfg, np.stack([s[:k] for s in states[:kt]])); cdt = time.perf_counter() - t0
eng.stream_reset()
same = []
for t in range(kt):
    r = eng.stream_step(states[t][:k]); same.append(float((r["iters"] == ref["iters"][t]).mean()))
warm = rows[1:]
print(json.dumps({"what": "warm-started streaming MPC, one GPU, host to host per tick", "robots": n, "ticks": ticks,
                  "first_tick_ms": rows[0]["ms"], "warm_tick_ms_median": float(np.median([r["ms"] for r in warm])),
                  "warm_robot_ticks_per_s": n / float(np.median([r["ms"] for r in warm])) * 1e3,
                  "cold_robot_ticks_per_s": n / cold_dt, "warm_mean_iters": float(np.mean([r["mean_iters"] for r in warm])),
                  "cold_mean_iters": float(cold["iters"].mean()),
                  "warm_mean_factorisations": float(np.mean([r["mean_factorisations"] for r in warm])),
                  "all_solved": all(r["solved"] == 1.0 for r in rows),
                  "cpu_oracle_warm": {"robot_ticks_per_s": k * kt / cdt, "threads": ob.max_threads(), "sample": f"{k} robots x {kt} ticks (first tick cold)"},
                  "parity_same_iters_per_tick": same, "per_tick": rows}))
