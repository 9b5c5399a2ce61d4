"""BASELINE config 4: H = 30, 8192 states (seed 1004), this rank's shard; host to host."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = pkg.config_default(); cfg.horizon = 30
eng = pkg.MpcEngine(cfg, 0)
st = pkg.generate_states(1004, 0, n)
eng.compute_grf_batch(st[:296])
ts = []
for rep in range(2):
    t0 = time.perf_counter(); eng.load_states(st); eng.build_qp(); t1 = time.perf_counter(); eng.solve(); out = eng.get_results(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1))
import oracle_binding as ob
k = 64
t0 = time.perf_counter(); ref = ob.mpc_compute_grf(cfg, st[:k]); cdt = time.perf_counter() - t0
den = np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
rel = np.linalg.norm(out["grf"][:k].astype(np.float64) - ref["grf"], axis=1) / den
print(json.dumps({"config": 4, "what": "long-horizon Go1 MPC H=30 (360 var), one GPU, host to host", "n": n,
                  "build_ms": 1e3 * ts[-1][0], "solve_ms": 1e3 * ts[-1][1], "solves_per_s": n / sum(ts[-1]),
                  "mean_iters": float(out["iters"].mean()), "mean_factorisations": 1 + float(out["rho_updates"].mean()),
                  "all_solved": bool((out["status"] == 1).all()),
                  "cpu_oracle": {"solves_per_s": k / cdt, "threads": ob.max_threads(), "sample": k},
                  "parity_max_rel_grf_err": float(rel.max()), "parity_same_iters": float((ref["iters"] == out["iters"][:k]).mean())}))
