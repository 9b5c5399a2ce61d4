import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
cfg = pkg.config_default(); cfg.horizon = 30; cfg.structured_solver = 3
e = pkg.MpcEngine(cfg, 0)
for n in (2048, 8192):
    st = pkg.generate_states(1004, 0, n)
    e.compute_grf_batch(st)
    t0 = time.perf_counter()
    for _ in range(3): r = e.compute_grf_batch(st)
    dt = (time.perf_counter() - t0) / 3
    print(f"[time] n={n}: {dt*1e3:.2f} ms -> {n/dt/1e3:.1f} k solves/s (mean iters {r['iters'].mean():.1f})", flush=True)
e.close()
