"""Pins the oracle's solver restatement with an INDEPENDENT second implementation (CPU only).

tests/osqp_independent.py is a numpy restatement of OSQP 0.6.x written from the paper: it
factorises the full quasi-definite KKT system with LAPACK LU every time rho changes, where the
oracle (and the GPU) work with the reduced SPD matrix K = P + sigma I + A' diag(rho) A.  The QP data
comes from the numpy restatement of ConvexMpc.cpp in tests/test_oracle.py, not from the oracle.
So everything from the robot record to the body-frame force is computed twice, by code that
shares nothing, and compared: iteration counts, rho updates, status, GRF.

Cases (VERDICT r01 "next round" 1a): 512 states x {gazebo, hardware} weights cold at the
benchmark tolerance, the reference's own library-default tolerance, an 8-tick warm-started stream
across a trot swap (osqp_update_P / _lin_cost / _bounds order), the stance-balance QP, H = 30.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

import osqp_independent as oi
from test_oracle import numpy_build

N_COLD = 512
TOL_GRF = 1e-3   # the GPU-vs-oracle gate of BASELINE.json, applied here oracle-vs-independent


def _bounds(cfg, rec):
    """ConvexMpc.cpp:223-245: per leg (0, -inf, 0, -inf, fz_min c) .. (inf, 0, inf, 0, fz_max c), replicated."""
    H = cfg.horizon
    l1, u1 = [], []
    for leg in range(4):
        c = float(rec["contacts"][leg] != 0)
        l1 += [0.0, -oi.OSQP_INFTY, 0.0, -oi.OSQP_INFTY, cfg.fz_min * c]
        u1 += [oi.OSQP_INFTY, 0.0, oi.OSQP_INFTY, 0.0, cfg.fz_max * c]
    return np.tile(l1, H), np.tile(u1, H)


def _cfg_by_name(name, H=10, eps=None):
    import go1_qp_mpc_controller_b200 as pkg
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    cfg.horizon = H
    if eps is not None:
        cfg.osqp.eps_abs = cfg.osqp.eps_rel = eps
    return cfg


def _cold_worker(args):
    name, H, eps, recs = args
    cfg = _cfg_by_name(name, H, eps)
    A = oi.mpc_constraint_matrix(H, cfg.mu)
    st = oi.Settings.from_ctypes(cfg.osqp)
    out = []
    for rec in recs:
        b = numpy_build(cfg, rec)
        l, u = _bounds(cfg, rec)
        s = oi.Osqp(b["P"], b["q"], A, l, u, st).solve()
        out.append((s.status, s.iters, s.rho_updates, oi.grf_body(s.solution(), rec["rot_mat"]), s.pri_res))
    return out


def _stream_worker(args):
    name, recs_t = args          # recs_t: (ticks, robots)
    cfg = _cfg_by_name(name)
    A = oi.mpc_constraint_matrix(10, cfg.mu)
    st = oi.Settings.from_ctypes(cfg.osqp)
    T, N = recs_t.shape
    out = np.zeros((T, N, 15))
    for i in range(N):
        solver = None
        for t in range(T):
            rec = recs_t[t, i]
            b = numpy_build(cfg, rec)
            l, u = _bounds(cfg, rec)
            if solver is None:                      # A1RobotControl.cpp:522-531 initSolver
                solver = oi.Osqp(b["P"], b["q"], A, l, u, st)
            else:                                   # :532-538 update*, warm start
                solver.update(b["P"], b["q"], l, u)
            solver.solve()
            out[t, i, :3] = (solver.status, solver.iters, solver.rho_updates)
            out[t, i, 3:] = oi.grf_body(solver.solution(), rec["rot_mat"])
    return out


def _pool_map(fn, jobs):
    # spawn, not fork: the parent has OpenMP (liboracle) and BLAS thread pools that do not survive a fork
    workers = min(len(jobs), max(1, len(os.sched_getaffinity(0))))
    old = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in old:
        os.environ[k] = "1"           # one problem per process; inherited by the spawned workers
    try:
        with mp.get_context("spawn").Pool(workers) as pool:
            return pool.map(fn, jobs)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _rel(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1.0)


@pytest.mark.parametrize("name,eps", [("gazebo", None), ("hardware", None), ("gazebo", 1e-3)])
def test_cold_solve_agrees_with_independent_implementation(pkg, ob, name, eps):
    cfg = _cfg_by_name(name, 10, eps)
    n = N_COLD if eps is None else 128
    states = pkg.generate_states(1002, 0, n)
    chunks = np.array_split(np.arange(n), 16)
    res = sum(_pool_map(_cold_worker, [(name, 10, eps, states[c]) for c in chunks]), [])
    ref = ob.mpc_compute_grf(cfg, states)
    status = np.array([r[0] for r in res])
    iters = np.array([r[1] for r in res])
    rho_up = np.array([r[2] for r in res])
    grf = np.stack([r[3] for r in res])
    assert np.array_equal(status, ref["status"]) and (status == 1).all()
    # two different factorisations of two different matrices: rounding may flip a termination check on
    # a rare state; everything else must walk the same iterate sequence
    same = iters == ref["iters"]
    assert same.mean() >= 0.995, same.mean()
    assert (rho_up[same] == ref["rho_updates"][same]).all()
    assert _rel(grf[same], ref["grf"][same]).max() <= 1e-6
    # ... and the states whose count differs are still inside the GRF gate? They stop 25 iterations
    # apart, so only the looser statement holds: both are eps-solutions of the same strongly convex QP
    if (~same).any():
        assert np.abs(iters[~same] - ref["iters"][~same]).max() <= 50
    pri = np.array([r[4] for r in res])
    assert np.median(np.abs(pri[same] - ref["pri_res"][same]) / np.maximum(ref["pri_res"][same], 1e-12)) < 1e-3


@pytest.mark.parametrize("name", ["gazebo", "hardware"])
def test_warm_stream_agrees_with_independent_implementation(pkg, ob, name):
    """8 control ticks of one persistent solver per robot across a trot contact swap (tick 48)."""
    cfg = _cfg_by_name(name)
    N, T = 64, 8
    st = np.stack([pkg.generate_stream_states(1006, 0, N, 44 + t) for t in range(T)])
    assert (st[3]["contacts"] != st[4]["contacts"]).any()       # the swap is inside the window
    ref = ob.mpc_stream(cfg, st)
    chunks = np.array_split(np.arange(N), 8)
    got = np.concatenate(_pool_map(_stream_worker, [(name, st[:, c]) for c in chunks]), axis=1)
    for t in range(T):
        assert np.array_equal(got[t, :, 0].astype(int), ref["status"][t]), t
        same = got[t, :, 1].astype(int) == ref["iters"][t]
        assert same.mean() >= 0.98, (t, same.mean())
        assert (got[t, same, 2].astype(int) == ref["rho_updates"][t][same]).all(), t
        assert _rel(got[t, same, 3:], ref["grf"][t][same]).max() <= 1e-5, t
        # a solver whose count differed on an earlier tick carries different iterates: skip it afterwards
        if not same.all():
            keep = same
            got, st = got[:, keep], st[:, keep]
            ref = ref[:, keep]
    assert got.shape[1] >= 0.9 * N


def test_long_horizon_agrees_with_independent_implementation(pkg, ob):
    cfg = _cfg_by_name("gazebo", 30)
    states = pkg.generate_states(1004, 0, 8)
    res = sum(_pool_map(_cold_worker, [("gazebo", 30, None, states[i:i + 1]) for i in range(8)]), [])
    ref = ob.mpc_compute_grf(cfg, states)
    assert np.array_equal([r[0] for r in res], ref["status"])
    assert np.array_equal([r[1] for r in res], ref["iters"])
    assert np.array_equal([r[2] for r in res], ref["rho_updates"])
    assert _rel(np.stack([r[3] for r in res]), ref["grf"]).max() <= 1e-6


def test_balance_qp_agrees_with_independent_implementation(pkg, ob):
    """Stance-balance QP (A1RobotControl.cpp:377-444): QP data from the oracle's build (that build is
    pinned by test_oracle.py::test_balance_oracle), solved by the independent implementation."""
    bcfg = pkg.balance_config_default()
    n = 256
    states = pkg.generate_balance_states(1005, 0, n)
    ref = ob.balance_compute_grf(bcfg, states)
    A = oi.balance_constraint_matrix(bcfg.mu)
    st = oi.Settings.from_ctypes(bcfg.osqp)
    same = 0
    for i in range(n):
        P, q, l, u = ob.balance_build_qp(bcfg, states[i])
        s = oi.Osqp(P, q, A, l, u, st).solve()
        assert s.status == ref["status"][i]
        if s.iters == ref["iters"][i]:
            same += 1
            assert s.rho_updates == ref["rho_updates"][i]
            g = oi.grf_body(s.solution(), states[i]["rot_mat"])
            assert _rel(g[None], ref["grf"][i][None]).max() <= 1e-6
    assert same >= 0.99 * n


def test_infeasible_problems_are_certified_alike(pkg, ob):
    """Primal-infeasible (contradictory bounds) and dual-infeasible (unbounded below) toy QPs on the MPC
    constraint matrix: the independent implementation and the oracle agree on OSQP's status codes."""
    cfg = _cfg_by_name("hardware")
    A = oi.mpc_constraint_matrix(10, cfg.mu)
    st = oi.Settings.from_ctypes(cfg.osqp)
    rec = pkg.generate_states(1002, 0, 1)[0]
    b = numpy_build(cfg, rec)
    l, u = _bounds(cfg, rec)
    # primal infeasible: fz >= 50 on a leg whose friction rows force fz <= ... : fx + mu fz in [0, inf),
    # fx - mu fz in (-inf, 0]  -> make row 4 demand fz in [-20, -10] (mu fz >= |fx| >= 0 contradicts)
    l2, u2 = l.copy(), u.copy()
    l2[4], u2[4] = -20.0, -10.0
    s = oi.Osqp(b["P"], b["q"], A, l2, u2, st).solve()
    x, y, info = ob.osqp_solve_mpc(cfg, b["P"], b["q"], l2, u2)
    assert s.status == oi.PRIMAL_INFEASIBLE and info["status"] == oi.PRIMAL_INFEASIBLE
    assert s.iters == info["iters"]
    assert np.isnan(s.solution()).all() and np.isnan(x).all()
    # dual infeasible: zero curvature and a descent direction along an unbounded ray
    P0 = np.zeros_like(b["P"])
    q0 = np.zeros_like(b["q"])
    q0[2] = -1.0                       # push fz of leg-step 0 up ...
    u3 = u.copy()
    u3[4] = oi.OSQP_INFTY              # ... with its upper bound removed
    s = oi.Osqp(P0, q0, A, l, u3, st).solve()
    x, y, info = ob.osqp_solve_mpc(cfg, P0, q0, l, u3)
    assert s.status == oi.DUAL_INFEASIBLE and info["status"] == oi.DUAL_INFEASIBLE
    assert s.iters == info["iters"]
