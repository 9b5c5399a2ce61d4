"""Regenerates tests/golden/*.npz from the oracle (run in the build container).

The reference stores no expected outputs for this path, so these
vectors pin (a) the oracle against regressions and (b) the GPU path against
committed numbers that travel to the GPU box.  Inputs come from the product's
counter-based generator; outputs from oracle/liboracle.so.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402


def stream():
    # --- warm-started streaming: 8 consecutive ticks across a trot swap (ticks 44..51) ---
    for name, cfg in (("gazebo", pkg.config_default()), ("hardware", pkg.config_hardware())):
        st = np.stack([pkg.generate_stream_states(1006, 0, 48, 44 + t) for t in range(8)])
        res = ob.mpc_stream(cfg, st)
        np.savez_compressed(os.path.join(HERE, f"stream_{name}.npz"), states=st, grf=res["grf"], iters=res["iters"],
                            status=res["status"], rho_updates=res["rho_updates"])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "stream":
        return stream()
    stream()
    # --- MPC, gazebo weights (primary) and hardware weights (secondary) ---
    for name, cfg in (("gazebo", pkg.config_default()), ("hardware", pkg.config_hardware())):
        states = pkg.generate_states(1002, 0, 96)
        res, sol = ob.mpc_compute_grf(cfg, states, want_solutions=True)
        P0, q0, l0, u0 = ob.mpc_build_qp(cfg, states[0])
        Psum = np.array([ob.mpc_build_qp(cfg, states[i])[0].sum() for i in range(8)])
        qs = np.stack([ob.mpc_build_qp(cfg, states[i])[1] for i in range(8)])
        np.savez_compressed(os.path.join(HERE, f"mpc_{name}.npz"), states=states, grf=res["grf"],
                            iters=res["iters"], status=res["status"], rho_updates=res["rho_updates"],
                            solutions=sol, P0=P0, q0=q0, l0=l0, u0=u0, Psum=Psum, q8=qs)
    # --- long horizon, H = 30 (BASELINE config 4) ---
    cfg = pkg.config_default()
    cfg.horizon = 30
    states = pkg.generate_states(1004, 0, 16)
    res = ob.mpc_compute_grf(cfg, states)
    np.savez_compressed(os.path.join(HERE, "mpc_h30.npz"), states=states, grf=res["grf"], iters=res["iters"],
                        status=res["status"], rho_updates=res["rho_updates"])
    # --- the reference's own driver input, test/test_mpc.cpp:15-60 ---
    cfg = pkg.config_default()
    cfg.mass = 15.0
    for i, v in enumerate([0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542]):
        cfg.inertia[i] = v
    for i, v in enumerate([1.0, 1.0, 1.0, 0.0, 0.0, 50.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0]):
        cfg.q_weights[i] = v
    for i in range(12):
        cfg.r_weights[i] = 1e-6
    st = np.zeros(1, dtype=pkg.abi.STATE_DTYPE)
    st["pos"][0] = [0, 0, 0.15]
    st["pos_d_z"][0] = 0.15      # test_mpc.cpp:83 uses root_pos[2] (+0) as the z reference
    st["rot_mat"][0] = np.eye(3).reshape(9)
    # test_mpc.cpp:105 feeds foot_pos_rel as the lever arms
    st["foot_pos_abs"][0] = [0.17, 0.15, -0.35, 0.17, -0.15, -0.35, -0.17, 0.15, -0.35, -0.17, -0.15, -0.35]
    st["contacts"][0] = [1, 0, 1, 0]
    out = {}
    for tag, eps in (("eps1e-5", 1e-5), ("eps1e-3", 1e-3), ("tight", 1e-10)):
        cfg.osqp.eps_abs = cfg.osqp.eps_rel = eps
        cfg.osqp.max_iter = 20000 if tag == "tight" else 4000
        r, s = ob.mpc_compute_grf(cfg, st, want_solutions=True)
        out[f"grf_{tag}"] = r["grf"][0]
        out[f"iters_{tag}"] = r["iters"][0]
    P, q, l, u = ob.mpc_build_qp(cfg, st[0])
    np.savez_compressed(os.path.join(HERE, "test_mpc_case.npz"), state=st, P=P, q=q, **out)
    # --- stance-balance QP ---
    bcfg = pkg.balance_config_default()
    bst = pkg.generate_balance_states(1005, 0, 128)
    bres = ob.balance_compute_grf(bcfg, bst)
    Pb, qb, lb, ub = ob.balance_build_qp(bcfg, bst[0])
    np.savez_compressed(os.path.join(HERE, "balance.npz"), states=bst, grf=bres["grf"], iters=bres["iters"],
                        status=bres["status"], P0=Pb, q0=qb, l0=lb, u0=ub)
    # --- generator known-answer: first records of three streams ---
    np.savez_compressed(os.path.join(HERE, "generator.npz"),
                        s1001=pkg.generate_states(1001, 0, 4), s1003_off=pkg.generate_states(1003, 65530, 4),
                        b1005=pkg.generate_balance_states(1005, 999_998, 2))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
