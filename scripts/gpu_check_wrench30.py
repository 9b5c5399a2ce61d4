"""GPU check of the long-horizon wrench-space engine (wrench_riccati_kernel.cuh): parity against the oracle
and against the Riccati engine (structured_solver = 1), warm stream, flags, timing.  Run under gpurun."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402

small = "--small" in sys.argv
timing_only = "--time" in sys.argv
N = 48 if small else 300


def rel(a, b):
    return np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1.0)


def mk(name, solver=0, flags=(0, 0, 0)):
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    cfg.horizon = 30
    cfg.structured_solver = solver
    cfg.exact_discretization, cfg.foot_drift, cfg.gait_aware = flags
    return cfg


if not timing_only:
    for name in ("gazebo", "hardware"):
        cfg = mk(name, 3)
        states = pkg.generate_states(1004, 0, N)
        e = pkg.MpcEngine(cfg, 0)
        res = e.compute_grf_batch(states).copy()
        ref = ob.mpc_compute_grf(cfg, states)
        print(f"[{name}] status", np.unique(res["status"], return_counts=True), flush=True)
        print(f"[{name}] same iters {(res['iters'] == ref['iters']).mean():.4f} same rho_updates "
              f"{(res['rho_updates'] == ref['rho_updates']).mean():.4f} max GRF rel {rel(res['grf'], ref['grf']).max():.2e} "
              f"mean iters {res['iters'].mean():.1f}", flush=True)
        bad = np.argsort(-rel(res["grf"], ref["grf"]))[:3]
        for i in bad:
            print("   worst", i, res["iters"][i], ref["iters"][i], res["rho_updates"][i], ref["rho_updates"][i],
                  res["grf"][i][:3], ref["grf"][i][:3])
        x = e.get_solution(0)
        P, q, l, u = e.get_qp(0)
        Po, qo, lo, uo = ob.mpc_build_qp(cfg, states[0])
        print(f"[{name}] get_qp on demand: P rel {np.abs(P - Po).max() / np.abs(Po).max():.2e}, solution finite {np.isfinite(x).all()}")
        if not small:
            d = pkg.MpcEngine(mk(name, 1), 0)
            rd = d.compute_grf_batch(states).copy()
            print(f"[{name}] vs Riccati engine: same iters {(res['iters'] == rd['iters']).mean():.4f} "
                  f"max GRF diff {np.abs(res['grf'] - rd['grf']).max():.3e}", flush=True)
            d.close()
        T = 3 if small else 6
        n = 16 if small else 96
        st = np.stack([pkg.generate_stream_states(1006, 0, n, 44 + t) for t in range(T)])
        sref = ob.mpc_stream(cfg, st)
        for t in range(T):
            r = e.stream_step(st[t])
            print(f"[{name}] warm tick {t}: same iters {(r['iters'] == sref['iters'][t]).mean():.4f} mean iters {r['iters'].mean():.1f} "
                  f"max GRF rel {rel(r['grf'], sref['grf'][t]).max():.2e}", flush=True)
        e.close()
    if not small:
        for flags in ((0, 1, 0), (0, 0, 1), (0, 1, 1)):
            cfg = mk("gazebo", 3, flags)
            n = 96
            st = pkg.generate_states(1002, 0, n)
            gait = pkg.generate_gait_inputs(1002, 0, n, 0)
            e = pkg.MpcEngine(cfg, 0)
            e.load_states(st)
            if cfg.gait_aware:
                e.set_gait_inputs(gait)
            e.build_qp()
            e.solve()
            r = e.get_results()
            ref = ob.mpc_compute_grf_ext(cfg, st, gait)
            print(f"[flags {flags}] same iters {(r['iters'] == ref['iters']).mean():.4f} max GRF rel {rel(r['grf'], ref['grf']).max():.2e}",
                  flush=True)
            e.close()

if not small:
    for mode, label in ((3, "wrench-riccati"), (1, "riccati")):
        e = pkg.MpcEngine(mk("gazebo", mode), 0)
        for n in (2048, 8192):
            st = pkg.generate_states(1004, 0, n)
            e.compute_grf_batch(st)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                r = e.compute_grf_batch(st)
            dt = (time.perf_counter() - t0) / reps
            print(f"[time] {label} n={n}: {dt * 1e3:.3f} ms per batch host-to-host -> {n / dt / 1e3:.1f} k solves/s "
                  f"(mean iters {r['iters'].mean():.1f})", flush=True)
        e.close()
print("done")
