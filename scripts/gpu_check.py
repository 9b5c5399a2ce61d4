"""First-light GPU check: QP-build parity, GRF parity and a rough timing.  Run under gpurun."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
import oracle_binding as ob

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = pkg.config_default()
states = pkg.generate_states(1002, 0, N)
eng = pkg.MpcEngine(cfg, 0)
eng.load_states(states)
eng.build_qp()
worstP = worstq = 0.0
for i in range(min(N, 16)):
    P, q, l, u = eng.get_qp(i)
    Po, qo, lo, uo = ob.mpc_build_qp(cfg, states[i])
    eP = np.abs(P - Po).max() / np.abs(Po).max(); eq = np.abs(q - qo).max() / np.abs(qo).max()
    worstP = max(worstP, eP); worstq = max(worstq, eq)
    assert np.array_equal(l, lo.astype(np.float32)) and np.array_equal(u, uo.astype(np.float32)), "bounds differ"
print(f"build parity: max rel P {worstP:.2e}  q {worstq:.2e}  bounds exact", flush=True)
t = time.time(); eng.solve(); print("solve wall", time.time() - t, flush=True)
res = eng.get_results()
ref = ob.mpc_compute_grf(cfg, states)
den = np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
rel = np.linalg.norm(res["grf"].astype(np.float64) - ref["grf"], axis=1) / den
print("status", np.unique(res["status"], return_counts=True))
print("iters gpu", np.unique(res["iters"], return_counts=True))
print("same iters", (res["iters"] == ref["iters"]).mean(), "same rho_updates", (res["rho_updates"] == ref["rho_updates"]).mean())
print(f"GRF rel err: max {rel.max():.2e} median {np.median(rel):.2e} frac>1e-3 {(rel>1e-3).mean():.4f} frac>1e-4 {(rel>1e-4).mean():.4f}")
bad = np.argsort(-rel)[:5]
for i in bad:
    print(i, rel[i], res["iters"][i], ref["iters"][i], res["rho_updates"][i], ref["rho_updates"][i], res["grf"][i][:6], ref["grf"][i][:6])
# timing
import ctypes
for n in (N, 4096):
    st = pkg.generate_states(1002, 0, n)
    eng.load_states(st); eng.synchronize()
    for rep in range(3):
        t0 = time.time(); eng.build_qp(); t1 = time.time(); eng.solve(); t2 = time.time()
        print(f"n={n}: build {1e3*(t1-t0):.2f} ms  solve {1e3*(t2-t1):.2f} ms  -> {n/(t2-t0):.0f} solves/s", flush=True)
    t0 = time.time(); out = eng.compute_grf_batch(st); t1 = time.time()
    print(f"n={n}: e2e {1e3*(t1-t0):.2f} ms -> {n/(t1-t0):.0f} solves/s  mean iters {out['iters'].mean():.1f} mean rho_updates {out['rho_updates'].mean():.2f}")
# facade
im = ob.mpc_build_intermediates(cfg, states[0])
Bl = np.tile(im["B_d"], (10, 1))
P2, q2, l2, u2 = eng.qp_mats_from_model(im["A_d"], Bl, im["x0"], im["x_ref"], (states["contacts"][0] != 0).astype(np.int32))
Po, qo, lo, uo = ob.mpc_build_qp(cfg, states[0])
print("facade build rel P", np.abs(P2 - Po).max() / np.abs(Po).max(), "q", np.abs(q2 - qo).max() / np.abs(qo).max())
x, status, iters = eng.solve_qp(Po, qo, lo, uo)
xo, yo, info = ob.osqp_solve_mpc(cfg, Po, qo, lo, uo)
print("facade solve: status", status, "iters", iters, info["iters"], "rel x[:12]", np.linalg.norm(x[:12] - xo[:12]) / max(np.linalg.norm(xo[:12]), 1))
# balance
bcfg = pkg.balance_config_default()
bst = pkg.generate_balance_states(1005, 0, 1024)
beng = pkg.MpcEngine(bcfg, 0, balance=True)
t0 = time.time(); bres = beng.compute_grf_batch(bst); t1 = time.time()
bref = ob.balance_compute_grf(bcfg, bst)
den = np.maximum(np.linalg.norm(bref["grf"], axis=1), 1.0)
brel = np.linalg.norm(bres["grf"].astype(np.float64) - bref["grf"], axis=1) / den
print("balance: status", np.unique(bres["status"], return_counts=True), "same iters", (bres["iters"] == bref["iters"]).mean(),
      f"rel max {brel.max():.2e} med {np.median(brel):.2e} frac>1e-3 {(brel>1e-3).mean():.4f}", "iters", bres["iters"].min(), bres["iters"].max(), "time ms", 1e3*(t1-t0))
Pb, qb, lb_, ub_ = beng.get_qp(0); Pbo, qbo, lbo, ubo = ob.balance_build_qp(bcfg, bst[0])
print("balance build rel P", np.abs(Pb - Pbo).max() / np.abs(Pbo).max(), "q", np.abs(qb - qbo).max() / np.abs(qbo).max(), "bounds", np.array_equal(lb_, lbo.astype(np.float32)), np.array_equal(ub_, ubo.astype(np.float32)))
