import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import numpy as np, go1_qp_mpc_controller_b200 as pkg
import oracle_binding as ob
def rel(a,b): return np.linalg.norm(a.astype(np.float64)-b,axis=1)/np.maximum(np.linalg.norm(b,axis=1),1.0)
for H in (10,30):
    cfg = pkg.config_default(); cfg.horizon = H
    st = pkg.generate_states(1004, 0, 32)
    fp = st["foot_pos_abs"].reshape(-1,4,3).copy()
    fp[:, :, :] = fp[:, :1, :]     # all four feet at ONE point: the wrench map has rank 3, N_k is singular
    st["foot_pos_abs"] = fp.reshape(st["foot_pos_abs"].shape)
    ref = ob.mpc_compute_grf(cfg, st)
    for solver in (0, 1 if H == 30 else 2):
        cfg.structured_solver = solver
        e = pkg.MpcEngine(cfg, 0)
        r = e.compute_grf_batch(st)
        print(f"H={H} solver {solver}: status {sorted(set(r['status'].tolist()))} same iters {(r['iters']==ref['iters']).mean():.3f} max GRF rel {rel(r['grf'],ref['grf']).max():.2e} oracle status {sorted(set(ref['status'].tolist()))} iters {ref['iters'][:6]} vs {r['iters'][:6]}")
        e.close()
