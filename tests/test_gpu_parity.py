"""GPU parity tests: the CUDA path through the C ABI against the oracle on the same seeded
inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes --
through size-independent properties.

Gates (BASELINE.json north_star / SURVEY.md 8d):
  QP matrices  max|dP| / max|P| <= 1e-5 (fp32 storage vs double), bounds exact
  GRF          ||f_gpu - f_oracle||_2 / max(||f_oracle||_2, 1 N) <= 1e-3, same solver settings
  constraints  friction-cone / bound violation of the returned x <= 1e-4 * fz_max
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOL_QP = 1e-5
TOL_GRF = 1e-3
TOL_CONE = 1e-4


# A state whose termination check flipped (rounding moved a residual across its tolerance) stops one
# check interval earlier or later than the oracle.  Both results are eps-solutions of the same QP, but at
# eps 1e-5 the first-step force is only determined to ~1e-2 in the weakly curved directions (P has 6 H
# eigenvalues equal to 2 r, SURVEY.md 7 hard part 1), so such a state is gated at TOL_GRF_FLIPPED, and
# the FRACTION of such states is gated separately (zero wherever that is what is measured).
TOL_GRF_FLIPPED = 5e-2


def grf_rel(gpu, ref):
    den = np.maximum(np.linalg.norm(ref, axis=1), 1.0)
    return np.linalg.norm(gpu.astype(np.float64) - ref, axis=1) / den


def assert_same_iterates(res, ref, max_flipped=0.0, what=""):
    """EVERY state is checked: identical iteration count -> GRF inside TOL_GRF; a flipped termination
    check (at most the fraction max_flipped of the states) -> counts one interval apart, GRF inside
    TOL_GRF_FLIPPED."""
    assert np.array_equal(res["status"], ref["status"]), what
    same = res["iters"] == ref["iters"]
    assert (~same).mean() <= max_flipped, (what, float((~same).mean()))
    rel = grf_rel(res["grf"], ref["grf"])
    assert rel[same].max(initial=0.0) <= TOL_GRF, (what, float(rel[same].max(initial=0.0)))
    assert (res["rho_updates"][same] == ref["rho_updates"][same]).all(), what
    if (~same).any():
        assert np.abs(res["iters"][~same] - ref["iters"][~same]).max() <= 50, what
        assert rel[~same].max() <= TOL_GRF_FLIPPED, (what, float(rel[~same].max()))


@pytest.fixture(scope="module")
def eng(pkg):
    e = pkg.MpcEngine(pkg.config_default(), 0)
    yield e
    e.close()


@pytest.mark.parametrize("which", ["gazebo", "hardware"])
def test_qp_build_parity(pkg, ob, which):
    cfg = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(1002, 0, 1024)
    e.load_states(states)
    e.build_qp()
    worstP = worstq = 0.0
    for i in range(0, 1024, 4):
        P, q, l, u = e.get_qp(i)
        Po, qo, lo, uo = ob.mpc_build_qp(cfg, states[i])
        worstP = max(worstP, np.abs(P - Po).max() / np.abs(Po).max())
        worstq = max(worstq, np.abs(q - qo).max() / np.abs(qo).max())
        assert np.array_equal(l, lo.astype(np.float32)) and np.array_equal(u, uo.astype(np.float32))
        assert np.abs(P - P.T).max() <= 2e-7 * np.abs(P).max()  # symmetric up to fp32 rounding
    assert worstP <= TOL_QP and worstq <= TOL_QP, (worstP, worstq)
    e.close()


@pytest.mark.parametrize("which,seed", [("gazebo", 1002), ("gazebo", 1003), ("hardware", 1002)])
def test_grf_parity_vs_oracle(pkg, ob, which, seed):
    cfg = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(seed, 0, 1024)
    res = e.compute_grf_batch(states)
    ref = ob.mpc_compute_grf(cfg, states)
    assert (res["status"] == 1).all() and (ref["status"] == 1).all()
    rel = grf_rel(res["grf"], ref["grf"])
    assert rel.max() <= TOL_GRF, (rel.max(), int(np.argmax(rel)))
    # same iterate sequence: identical iteration and refactorisation counts
    assert (res["iters"] == ref["iters"]).mean() >= 0.995
    assert (res["rho_updates"] == ref["rho_updates"]).mean() >= 0.995
    # the exit residual is a difference of nearly equal numbers: compare it statistically
    rel_res = np.abs(res["pri_res"] - ref["pri_res"]) / np.maximum(ref["pri_res"], 1e-9)
    assert np.median(rel_res) < 0.01 and (rel_res < 0.5).mean() >= 0.98
    e.close()


def test_reference_default_tolerances(pkg, ob):
    """The reference itself runs OSQP at eps 1e-3 (A1RobotControl.cpp:523-524 changes nothing else)."""
    cfg = pkg.config_default()
    d = pkg.settings_osqp_default()
    cfg.osqp.eps_abs, cfg.osqp.eps_rel = d.eps_abs, d.eps_rel
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(1002, 2000, 512)
    res = e.compute_grf_batch(states)
    ref = ob.mpc_compute_grf(cfg, states)
    assert (res["iters"] == ref["iters"]).mean() >= 0.995
    assert grf_rel(res["grf"], ref["grf"]).max() <= TOL_GRF
    e.close()


@pytest.mark.parametrize("name", ["gazebo", "hardware"])
def test_golden_fixtures(pkg, name):
    g = np.load(os.path.join(GOLD, f"mpc_{name}.npz"))
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    e = pkg.MpcEngine(cfg, 0)
    e.load_states(g["states"])
    e.build_qp()
    P, q, l, u = e.get_qp(0)
    assert np.abs(P - g["P0"]).max() / np.abs(g["P0"]).max() <= TOL_QP
    assert np.abs(q - g["q0"]).max() / np.abs(g["q0"]).max() <= TOL_QP
    assert np.array_equal(l, g["l0"].astype(np.float32)) and np.array_equal(u, g["u0"].astype(np.float32))
    for i in range(8):
        Pi, qi, _, _ = e.get_qp(i)
        assert abs(Pi.astype(np.float64).sum() - g["Psum"][i]) <= 1e-5 * np.abs(g["P0"]).max() * 120
        assert np.abs(qi - g["q8"][i]).max() <= TOL_QP * np.abs(g["q8"][i]).max()
    e.solve()
    res = e.get_results()
    assert np.array_equal(res["iters"], g["iters"]) and np.array_equal(res["status"], g["status"])
    assert np.array_equal(res["rho_updates"], g["rho_updates"])
    assert grf_rel(res["grf"], g["grf"]).max() <= TOL_GRF
    x = np.stack([e.get_solution(i) for i in range(16)])
    den = np.maximum(np.linalg.norm(g["solutions"][:16], axis=1), 1.0)
    assert (np.linalg.norm(x - g["solutions"][:16], axis=1) / den).max() <= TOL_GRF
    e.close()


def test_cone_and_bound_violation(pkg, eng):
    """Returned x (world frame, all H steps): |fx|,|fy| <= mu fz, 0 <= fz <= 180 c, to 1e-4 * fz_max."""
    states = pkg.generate_states(1004, 0, 512)
    eng.compute_grf_batch(states)
    worst = 0.0
    for i in range(0, 512, 8):
        x = eng.get_solution(i).astype(np.float64).reshape(10, 4, 3)
        c = states["contacts"][i].astype(np.float64)
        v = np.maximum(np.abs(x[..., :2]).max(-1) - 0.3 * x[..., 2], 0).max()
        v = max(v, np.maximum(-x[..., 2], 0).max(), np.maximum(x[..., 2] - 180.0 * c[None, :], 0).max())
        worst = max(worst, v)
    assert worst <= TOL_CONE * 180.0, worst


def test_convexmpc_surface_like_test_mpc_cpp(pkg, ob):
    """The reference's own driver (test/test_mpc.cpp:14-159) restated against this surface."""
    from go1_qp_mpc_controller_b200 import A1CtrlStates, ConvexMpc
    g = np.load(os.path.join(GOLD, "test_mpc_case.npz"))
    state = A1CtrlStates()
    state.robot_mass = 15
    state.a1_trunk_inertia = np.diag([0.0158533, 0.0377999, 0.0456542])
    state.root_euler = np.zeros(3)
    state.root_rot_mat = np.eye(3)
    state.root_pos = np.array([0.0, 0.0, 0.15])
    state.foot_pos_rel = np.array([[0.17, 0.17, -0.17, -0.17], [0.15, -0.15, 0.15, -0.15], [-0.35] * 4])
    state.contacts = [True, False, True, False]
    dt = 0.0025
    q_weights = np.array([1.0, 1.0, 1.0, 0.0, 0.0, 50.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0])
    r_weights = np.full(12, 1e-6)
    mpc_solver = ConvexMpc(q_weights, r_weights)
    mpc_solver.reset()
    state.mpc_states = np.concatenate([state.root_euler, state.root_pos, state.root_ang_vel,
                                       state.root_lin_vel, [-9.8]])
    state.root_lin_vel_d_world = state.root_rot_mat @ state.root_lin_vel_d
    xr = np.zeros(130)
    for i in range(10):  # test_mpc.cpp:77-92 (zero commands -> constant reference)
        xr[13 * i:13 * i + 13] = [0, 0, 0, 0, 0, state.root_pos[2], 0, 0, 0, 0, 0, 0, -9.8]
    state.mpc_states_d = xr
    mpc_solver.calculate_A_mat_c(np.zeros(3))
    foot = state.foot_pos_rel.copy()
    for i in range(10):
        mpc_solver.calculate_B_mat_c(state.robot_mass, state.a1_trunk_inertia, state.root_rot_mat, foot)
        mpc_solver.state_space_discretization(dt)
        mpc_solver.B_mat_d_list[13 * i:13 * i + 13, :] = mpc_solver.B_mat_d
    mpc_solver.calculate_qp_mats(state)
    assert np.abs(mpc_solver.hessian - g["P"]).max() / np.abs(g["P"]).max() <= TOL_QP
    assert np.abs(mpc_solver.gradient - g["q"]).max() / np.abs(g["q"]).max() <= TOL_QP
    assert mpc_solver.ub[4] == 180.0 and mpc_solver.ub[9] == 0.0 and mpc_solver.lb[1] == -1e30
    # OsqpEigen initSolver + solve + getSolution (test_mpc.cpp:131-151), cold start
    sol, status, iters = mpc_solver._engine.solve_qp(mpc_solver.hessian, mpc_solver.gradient,
                                                     mpc_solver.lb, mpc_solver.ub)
    assert status == 1 and iters == int(g["iters_eps1e-5"])
    grf = sol[:12]                         # world frame == body frame here (R = I)
    assert np.linalg.norm(grf - g["grf_eps1e-5"]) / np.linalg.norm(g["grf_eps1e-5"]) <= TOL_GRF
    np.testing.assert_allclose(grf[[1, 2]], [-12.782, 42.606], atol=5e-3)   # SURVEY.md 8c probe


def test_compute_grf_single_robot(pkg, ob):
    """A1RobotControl::compute_grf, both branches, one robot (BASELINE config 1)."""
    from go1_qp_mpc_controller_b200 import A1CtrlStates, A1RobotControl
    cfg = pkg.config_default()
    rec = pkg.generate_states(1001, 5, 1)
    st = A1CtrlStates()
    st.robot_mass = cfg.mass
    st.q_weights = np.array(cfg.q_weights[:])
    st.r_weights = np.array(cfg.r_weights[:])
    st.root_euler = rec["euler"][0].astype(np.float64)
    st.root_pos = rec["pos"][0].astype(np.float64)
    st.root_ang_vel = rec["ang_vel"][0].astype(np.float64)
    st.root_lin_vel = rec["lin_vel"][0].astype(np.float64)
    st.root_euler_d = rec["euler_d"][0].astype(np.float64)
    st.root_pos_d = np.array([0.0, 0.0, float(rec["pos_d_z"][0])])
    st.root_lin_vel_d = rec["lin_vel_d"][0].astype(np.float64)
    st.root_ang_vel_d = rec["ang_vel_d"][0].astype(np.float64)
    st.root_rot_mat = rec["rot_mat"][0].astype(np.float64).reshape(3, 3)
    st.foot_pos_abs = rec["foot_pos_abs"][0].astype(np.float64).reshape(4, 3).T
    st.contacts = [bool(c) for c in rec["contacts"][0]]
    ctl = A1RobotControl(cfg)
    grf = ctl.compute_grf(st, 0.0025)
    ref = ob.mpc_compute_grf(cfg, rec)["grf"][0].reshape(4, 3).T
    assert grf.shape == (3, 4)
    assert np.linalg.norm(grf - ref) / max(np.linalg.norm(ref), 1.0) <= TOL_GRF
    assert st.mpc_states[12] == -9.8
    # QP branch
    brec = pkg.generate_balance_states(1005, 11, 1)
    bcfg = pkg.balance_config_default()
    st.stance_leg_control_type = 0
    st.robot_mass = bcfg.mass
    st.kp_linear, st.kd_linear = np.array(bcfg.kp_linear[:]), np.array(bcfg.kd_linear[:])
    st.kp_angular, st.kd_angular = np.array(bcfg.kp_angular[:]), np.array(bcfg.kd_angular[:])
    for f, k in (("root_euler", "euler"), ("root_pos", "pos"), ("root_ang_vel", "ang_vel"),
                 ("root_lin_vel", "lin_vel"), ("root_euler_d", "euler_d"), ("root_pos_d", "pos_d"),
                 ("root_lin_vel_d", "lin_vel_d"), ("root_ang_vel_d", "ang_vel_d")):
        setattr(st, f, brec[k][0].astype(np.float64))
    st.root_rot_mat = brec["rot_mat"][0].astype(np.float64).reshape(3, 3)
    st.root_rot_mat_z = brec["rot_mat_z"][0].astype(np.float64).reshape(3, 3)
    st.foot_pos_abs = brec["foot_pos_abs"][0].astype(np.float64).reshape(4, 3).T
    st.contacts = [bool(c) for c in brec["contacts"][0]]
    grf = ctl.compute_grf(st, 0.0025)
    ref = ob.balance_compute_grf(bcfg, brec)["grf"][0].reshape(4, 3).T
    assert np.linalg.norm(grf - ref) / max(np.linalg.norm(ref), 1.0) <= TOL_GRF


def test_edge_cases(pkg, ob, eng):
    cfg = pkg.config_default()
    # empty batch
    out = eng.compute_grf_batch(np.zeros(0, dtype=pkg.abi.STATE_DTYPE))
    assert len(out) == 0
    # one state, and ragged sizes around the SM count (148) and the warp size
    for n in (1, 31, 147, 149, 297):
        states = pkg.generate_states(1002, 5000, n)
        res = eng.compute_grf_batch(states)
        ref = ob.mpc_compute_grf(cfg, states)
        assert np.array_equal(res["iters"], ref["iters"])
        assert grf_rel(res["grf"], ref["grf"]).max() <= TOL_GRF
    # every leg in swing: all fz rows are equalities [0,0] -> zero force
    states = pkg.generate_states(1002, 0, 8)
    states["contacts"][:] = 0.0
    res = eng.compute_grf_batch(states)
    ref = ob.mpc_compute_grf(cfg, states)
    assert np.abs(res["grf"]).max() < 1e-3 and np.abs(ref["grf"]).max() < 1e-3
    assert np.array_equal(res["status"], ref["status"]) and np.array_equal(res["iters"], ref["iters"])
    # single stance leg, three-leg stance: contact patterns the generator never draws
    for pat in ([1, 0, 0, 0], [1, 1, 1, 0], [0, 0, 1, 1]):
        states = pkg.generate_states(1002, 40, 16)
        states["contacts"][:] = pat
        res = eng.compute_grf_batch(states)
        ref = ob.mpc_compute_grf(cfg, states)
        assert_same_iterates(res, ref, max_flipped=0.07, what=str(pat))   # at most one state of the 16


def test_max_iter_status(pkg, ob):
    """Hitting max_iter reports OSQP's codes (-2, or 2 = solved inaccurate), like the oracle."""
    cfg = pkg.config_default()
    cfg.osqp.max_iter = 60
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(1002, 0, 128)
    res = e.compute_grf_batch(states)
    ref = ob.mpc_compute_grf(cfg, states)
    assert np.array_equal(res["status"], ref["status"])
    assert set(np.unique(res["status"])) <= {1, 2, -2}
    assert (res["iters"] <= 60).all() and np.array_equal(res["iters"], ref["iters"])
    e.close()


def test_error_behaviour(pkg):
    cfg = pkg.config_default()
    e = pkg.MpcEngine(cfg, 0)
    with pytest.raises(pkg.MpcError) as ei:
        e.solve()
    assert ei.value.code == pkg.abi.MPC_ERR_STATE
    with pytest.raises(pkg.MpcError):
        e.get_results()
    e.load_states(pkg.generate_states(1, 0, 4))
    e.build_qp()
    with pytest.raises(pkg.MpcError):
        e.get_qp(4)
    e.close()
    cfg.horizon = 20   # kernels exist for H = 10 and H = 30
    with pytest.raises(pkg.MpcError) as ei:
        pkg.MpcEngine(cfg, 0)
    assert ei.value.code == pkg.abi.MPC_ERR_UNSUPPORTED
    # caller QP with crossed bounds: refused like osqp_setup refuses it (the device evaluates no certificates)
    e = pkg.MpcEngine(pkg.config_default(), 0)
    n, m = 120, 200
    lb, ub = np.zeros(m), np.ones(m)
    lb[7] = 2.0
    with pytest.raises(pkg.MpcError) as ei:
        e.solve_qp(np.eye(n), np.zeros(n), lb, ub)
    assert ei.value.code == pkg.abi.MPC_ERR_INVALID
    e.close()
    cfg = pkg.config_default()
    cfg.osqp.adaptive_rho_interval = 0   # the wall-clock dependent library default is refused
    with pytest.raises(pkg.MpcError) as ei:
        pkg.MpcEngine(cfg, 0)
    assert ei.value.code == pkg.abi.MPC_ERR_INVALID


def test_full_size_properties(pkg, ob, eng):
    """BASELINE configs 2 and 3 (4096 and 65536 states): determinism, permutation invariance,
    feasibility, and a sampled oracle comparison."""
    cfg = pkg.config_default()
    states = pkg.generate_states(1002, 0, 4096)
    a = eng.compute_grf_batch(states).copy()
    b = eng.compute_grf_batch(states).copy()
    assert a.tobytes() == b.tobytes()                       # bit-identical repeat
    perm = np.random.default_rng(0).permutation(4096)
    c = eng.compute_grf_batch(states[perm]).copy()
    assert c.tobytes() == a[perm].tobytes()                 # a result depends on its own state only
    big = pkg.generate_states(1003, 0, 65536)
    r = eng.compute_grf_batch(big)
    assert (r["status"] == 1).all()
    assert r["iters"].min() >= 50 and r["iters"].max() <= 1000
    # body-frame forces rotate back to world-frame forces that respect fz in [0, 180 c]
    Rm = big["rot_mat"].reshape(-1, 3, 3).astype(np.float64)
    fw = np.einsum("nij,nlj->nli", Rm, r["grf"].reshape(-1, 4, 3).astype(np.float64))
    cmask = big["contacts"].astype(np.float64)
    assert (fw[..., 2] >= -TOL_CONE * 180).all() and (fw[..., 2] <= 180.0 * cmask + TOL_CONE * 180 + 2e-3).all()
    assert (np.abs(fw[..., :2]).max(-1) <= 0.3 * fw[..., 2] + TOL_CONE * 180 + 2e-3).all()
    idx = np.arange(0, 65536, 128)
    ref = ob.mpc_compute_grf(cfg, big[idx])
    assert (r["iters"][idx] == ref["iters"]).mean() >= 0.995
    assert grf_rel(r["grf"][idx], ref["grf"]).max() <= TOL_GRF
    # the first 4096 of a shard of the 65536 batch equal the stand-alone solve (sharding = slicing)
    lo = 8192
    s = eng.compute_grf_batch(big[lo:lo + 512])
    assert s.tobytes() == r[lo:lo + 512].tobytes()


def test_device_resident_path_and_external_stream(pkg, ob):
    """Inputs already in HBM (torch tensor) and kernels on a caller-owned stream."""
    import torch
    cfg = pkg.config_default()
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(1002, 123, 300)
    t = torch.from_numpy(states.view(np.uint8).reshape(-1).copy()).cuda()
    s = torch.cuda.Stream()
    e.set_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        e.set_states_device(t.data_ptr(), 300)
        e.build_qp(sync=False)
        e.solve(sync=False)
    s.synchronize()
    res = e.get_results()
    e.set_stream(0)
    ref = e.compute_grf_batch(states)
    assert res.tobytes() == ref.tobytes()
    assert e.kernel_launches() == 2   # one fused kernel per batch (build + solve)
    e.close()


def test_balance_qp_parity(pkg, ob):
    bcfg = pkg.balance_config_default()
    be = pkg.MpcEngine(bcfg, 0, balance=True)
    g = np.load(os.path.join(GOLD, "balance.npz"))
    res = be.compute_grf_batch(g["states"])
    assert np.array_equal(res["iters"], g["iters"]) and np.array_equal(res["status"], g["status"])
    assert grf_rel(res["grf"], g["grf"]).max() <= TOL_GRF
    P, q, l, u = be.get_qp(0)
    assert np.abs(P - g["P0"]).max() / np.abs(g["P0"]).max() <= TOL_QP
    assert np.abs(q - g["q0"]).max() / np.abs(g["q0"]).max() <= TOL_QP
    assert np.array_equal(l, g["l0"].astype(np.float32)) and np.array_equal(u, g["u0"].astype(np.float32))
    states = pkg.generate_balance_states(1005, 0, 8192)
    res = be.compute_grf_batch(states)
    ref = ob.balance_compute_grf(bcfg, states)
    assert_same_iterates(res, ref, max_flipped=0.005, what="balance 8192")
    # ragged / empty
    assert len(be.compute_grf_batch(np.zeros(0, dtype=pkg.abi.BALANCE_DTYPE))) == 0
    r1 = be.compute_grf_batch(states[:5])
    assert r1.tobytes() == res[:5].tobytes()
    be.close()


def test_long_horizon_h30(pkg, ob):
    """BASELINE config 4: H = 30 (360 variables, 600 constraints) through the generic-horizon kernels."""
    cfg = pkg.config_default()
    cfg.horizon = 30
    e = pkg.MpcEngine(cfg, 0)
    states = pkg.generate_states(1004, 0, 200)
    e.load_states(states)
    e.build_qp()
    for i in (0, 7, 199):
        P, q, l, u = e.get_qp(i)
        Po, qo, lo, uo = ob.mpc_build_qp(cfg, states[i])
        assert P.shape == (360, 360) and l.shape == (600,)
        assert np.abs(P - Po).max() / np.abs(Po).max() <= TOL_QP
        assert np.abs(q - qo).max() / np.abs(qo).max() <= TOL_QP
        assert np.array_equal(l, lo.astype(np.float32)) and np.array_equal(u, uo.astype(np.float32))
        assert np.array_equal(P, P.T)
    e.solve()
    res = e.get_results()
    ref = ob.mpc_compute_grf(cfg, states)
    assert (res["status"] == 1).all()
    assert_same_iterates(res, ref, max_flipped=0.0, what="H=30")
    x = e.get_solution(3).astype(np.float64).reshape(30, 4, 3)
    c = states["contacts"][3].astype(np.float64)
    assert np.maximum(np.abs(x[..., :2]).max(-1) - 0.3 * x[..., 2], 0).max() <= TOL_CONE * 180.0
    assert np.maximum(x[..., 2] - 180.0 * c[None, :], 0).max() <= TOL_CONE * 180.0
    g = np.load(os.path.join(GOLD, "mpc_h30.npz"))                 # committed fixture
    rg = e.compute_grf_batch(g["states"])
    assert np.array_equal(rg["iters"], g["iters"]) and np.array_equal(rg["rho_updates"], g["rho_updates"])
    assert grf_rel(rg["grf"], g["grf"]).max() <= TOL_GRF
    # ragged sizes around the SM count, and a repeat is bit-identical
    r2 = e.compute_grf_batch(states[:149])
    assert r2.tobytes() == res[:149].tobytes()
    assert len(e.compute_grf_batch(np.zeros(0, dtype=pkg.abi.STATE_DTYPE))) == 0
    e.close()


@pytest.mark.parametrize("which", ["gazebo", "hardware"])
def test_warm_started_stream_parity(pkg, ob, which):
    """The solver kept alive between control ticks (A1RobotControl.cpp:522-538): the GPU slots
    follow the oracle's OSQP update semantics tick by tick -- same iteration counts, same rho
    updates, GRF inside the gate -- across a trot contact swap (bounds change type)."""
    cfg = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
    N, T = 256, 8
    st = np.stack([pkg.generate_stream_states(1006, 0, N, 44 + t) for t in range(T)])
    ref = ob.mpc_stream(cfg, st)
    e = pkg.MpcEngine(cfg, 0)
    cold = e.compute_grf_batch(st[1]).copy()
    for t in range(T):
        res = e.stream_step(st[t])
        assert_same_iterates(res, ref[t], max_flipped=0.0, what=f"warm tick {t}")
        if t == 0:
            first = res.copy()
    # a warm tick is cheaper than the cold solve of the same problem
    res1 = None
    e.stream_reset()
    e.stream_step(st[0])
    res1 = e.stream_step(st[1])
    assert res1["iters"].mean() < 0.7 * cold["iters"].mean()
    # reset makes the next tick an initSolver again: identical to the plain cold path
    e.stream_reset()
    again = e.stream_step(st[0])
    plain = e.compute_grf_batch(st[0])
    assert np.array_equal(again["iters"], plain["iters"]) and np.array_equal(again["grf"], plain["grf"])
    assert np.array_equal(again["grf"], first["grf"])
    e.close()


@pytest.mark.parametrize("name", ["gazebo", "hardware"])
def test_warm_stream_golden(pkg, name):
    g = np.load(os.path.join(GOLD, f"stream_{name}.npz"))
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    e = pkg.MpcEngine(cfg, 0)
    for t in range(g["states"].shape[0]):
        res = e.stream_step(g["states"][t])
        assert np.array_equal(res["iters"], g["iters"][t]), t
        assert grf_rel(res["grf"], g["grf"][t]).max() <= TOL_GRF, t
    e.close()


def test_warm_stream_long_horizon(pkg, ob):
    """The persistent solver at H = 30 (Riccati-structured kernel): the same OSQP update semantics,
    tick by tick against the oracle's MpcStream, across a trot swap."""
    cfg = pkg.config_default()
    cfg.horizon = 30
    N, T = 48, 6
    st = np.stack([pkg.generate_stream_states(1006, 0, N, 45 + t) for t in range(T)])
    assert (st[2]["contacts"] != st[3]["contacts"]).any()
    ref = ob.mpc_stream(cfg, st)
    e = pkg.MpcEngine(cfg, 0)
    cold = e.compute_grf_batch(st[1]).copy()
    for t in range(T):
        res = e.stream_step(st[t])
        assert_same_iterates(res, ref[t], max_flipped=0.0, what=f"H=30 warm tick {t}")
    assert ref["iters"][1].mean() < 0.8 * cold["iters"].mean()      # a warm tick is cheaper
    e.stream_reset()
    again = e.stream_step(st[0])
    assert np.array_equal(again["iters"], ref["iters"][0])
    # the dense long-horizon workspace path has no warm start
    cfg2 = pkg.config_default()
    cfg2.horizon = 30
    cfg2.structured_solver = 2
    e2 = pkg.MpcEngine(cfg2, 0)
    e2.load_states(st[0][:2])
    e2.build_qp()
    with pytest.raises(pkg.MpcError) as ei:
        e2.solve_warm()
    assert ei.value.code == pkg.abi.MPC_ERR_UNSUPPORTED
    e.close()
    e2.close()


def _check_torques(ob, states, tin, res, tq, balance=False):
    # (a) the map itself: oracle map applied to the GPU's own GRF (fp32 output rounding only)
    tau_a, mask_a = ob.torque_map(states, tin, res["grf"].astype(np.float64), balance=balance)
    assert np.array_equal(tq["nan_mask"], mask_a)
    ok = ~np.isnan(tau_a)
    assert np.abs(tq["joint_torques"][ok] - tau_a[ok]).max() <= 1e-5 * max(1.0, np.abs(tau_a[ok]).max())
    return tau_a


@pytest.mark.parametrize("path", ["cold", "warm", "h30", "balance"])
def test_torque_map_parity(pkg, ob, path):
    """compute_joint_torques (A1RobotControl.cpp:289-319) fused into the result writer."""
    n = 96 if path == "h30" else 512
    tin = pkg.generate_torque_inputs(1002, 0, n)
    if path == "balance":
        cfg = pkg.balance_config_default()
        e = pkg.MpcEngine(cfg, 0, balance=True)
        st = pkg.generate_balance_states(1005, 0, n)
        e.load_states(st)
        e.set_torque_inputs(tin)
        e.solve()
        res, tq = e.get_results(), e.get_torques()
        ref = ob.balance_compute_grf(cfg, st)
        _check_torques(ob, st, tin, res, tq, balance=True)
        tau_ref, _ = ob.torque_map(st, tin, ref["grf"], balance=True)
    else:
        cfg = pkg.config_default()
        if path == "h30":
            cfg.horizon = 30
        e = pkg.MpcEngine(cfg, 0)
        st = pkg.generate_states(1002, 0, n)
        e.load_states(st)
        e.set_torque_inputs(tin)
        e.build_qp()
        if path == "warm":
            e.solve_warm()
        else:
            e.solve()
        res, tq = e.get_results(), e.get_torques()
        ref = ob.mpc_compute_grf(cfg, st)
        _check_torques(ob, st, tin, res, tq)
        tau_ref, _ = ob.torque_map(st, tin, ref["grf"])
    # (b) end to end against the oracle's GRF -> oracle map: the GRF gate carried through J
    assert np.array_equal(res["iters"], ref["iters"])
    err = np.abs(tq["joint_torques"] - tau_ref).max(axis=1) / np.maximum(np.abs(tau_ref).max(axis=1), 1.0)
    assert err.max() <= TOL_GRF, err.max()
    e.close()


def test_torque_map_nan_guard_and_errors(pkg, ob):
    n = 64
    e = pkg.MpcEngine(pkg.config_default(), 0)
    st = pkg.generate_states(1002, 0, n)
    tin = pkg.generate_torque_inputs(1002, 0, n)
    i, leg = np.argwhere(st["contacts"] == 0)[0]
    tin["j_foot"][i][9 * leg:9 * leg + 9] = 0.0          # singular swing-leg Jacobian -> NaN torques
    with pytest.raises(pkg.MpcError) as ei:
        e.set_torque_inputs(tin)                          # states first
    assert ei.value.code == pkg.abi.MPC_ERR_STATE
    e.load_states(st)
    e.build_qp()
    e.solve()
    with pytest.raises(pkg.MpcError):
        e.get_torques()                                   # no torque inputs were given
    e.load_states(st)
    e.set_torque_inputs(tin)
    e.build_qp()
    e.solve()
    tq = e.get_torques()
    assert (tq["nan_mask"][i] >> (3 * leg)) & 7 == 7
    assert (np.delete(tq["nan_mask"], i) == 0).all()
    assert (tq["joint_torques"][i][3 * leg:3 * leg + 3] == 0).all()
    _check_torques(ob, st, tin, e.get_results(), tq)
    with pytest.raises(pkg.MpcError):
        e.set_torque_inputs(tin[:5])                      # wrong count
    e.close()


def test_mirror_compute_joint_torques(pkg, ob):
    """Host mirror: zero torques for the first nine calls, then the device map (:292-295)."""
    ctl = pkg.A1RobotControl()
    s = pkg.A1CtrlStates()
    rec = pkg.generate_states(1002, 3, 1)
    tin = pkg.generate_torque_inputs(1002, 3, 1)
    s.root_euler, s.root_pos = rec["euler"][0].astype(float), rec["pos"][0].astype(float)
    s.root_ang_vel, s.root_lin_vel = rec["ang_vel"][0].astype(float), rec["lin_vel"][0].astype(float)
    s.root_euler_d = rec["euler_d"][0].astype(float)
    s.root_pos_d = np.array([0.0, 0.0, float(rec["pos_d_z"][0])])
    s.root_lin_vel_d, s.root_ang_vel_d = rec["lin_vel_d"][0].astype(float), rec["ang_vel_d"][0].astype(float)
    s.root_rot_mat = rec["rot_mat"][0].astype(float).reshape(3, 3)
    s.foot_pos_abs = rec["foot_pos_abs"][0].astype(float).reshape(4, 3).T
    s.contacts = [bool(c) for c in rec["contacts"][0]]
    for leg in range(4):
        s.j_foot[3 * leg:3 * leg + 3, 3 * leg:3 * leg + 3] = tin["j_foot"][0][9 * leg:9 * leg + 9].reshape(3, 3)
    s.foot_forces_kin = tin["foot_forces_kin"][0].astype(float).reshape(4, 3).T
    grf = ctl.compute_grf(s, 0.0025)
    for k in range(9):
        ctl.compute_joint_torques(s)
        assert (s.joint_torques == 0).all()
    ctl.compute_joint_torques(s)
    tau, mask = ob.torque_map(s.to_record(), s.to_torque_record(), grf.T.reshape(1, 12))
    assert mask[0] == 0 and np.allclose(s.joint_torques, tau[0], rtol=1e-5, atol=1e-5)


def test_state_preparation_parity(pkg, ob):
    """Orientation, leg kinematics, EKF and terrain pitch on the device against the oracle, tick by
    tick with persistent per-robot filters; then the prepared batch goes straight into the solver."""
    cfg = pkg.prep_config_default()
    N, T = 300, 14
    sens = np.stack([pkg.generate_sensors(1007, 0, N, 3 * t) for t in range(T)])
    st_o, tin_o, ex_o = ob.prep_stream(cfg, sens)
    e = pkg.MpcEngine(pkg.config_default(), 0)
    for t in range(T):
        e.prepare_states(sens[t], cfg)
        st, tin, ex = e.get_prepared()
        for rec, ref in ((st, st_o[t]), (tin, tin_o[t]), (ex, ex_o[t])):
            for f in rec.dtype.names:
                a, b = rec[f].astype(np.float64), ref[f].astype(np.float64)
                tol = 2e-5 if f.startswith("estimated_root") or f in ("pos", "lin_vel", "foot_pos_world", "foot_vel_world") else 2e-6
                assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (t, f, np.abs(a - b).max())
    # the prepared batch feeds the solver without leaving the device: same results as the host path
    e.build_qp()
    e.solve()
    res, tq = e.get_results(), e.get_torques()
    e2 = pkg.MpcEngine(pkg.config_default(), 0)
    e2.load_states(st)
    e2.set_torque_inputs(tin)
    e2.build_qp()
    e2.solve()
    res2, tq2 = e2.get_results(), e2.get_torques()
    assert np.array_equal(res["grf"], res2["grf"]) and np.array_equal(res["iters"], res2["iters"])
    assert np.array_equal(tq["joint_torques"], tq2["joint_torques"])
    assert (res["status"] == 1).all()
    # reset forgets the filters: the next tick initialises again (root_pos = odometry)
    e.prepare_reset()
    e.prepare_states(sens[5], cfg)
    st_r, _, ex_r = e.get_prepared()
    assert np.array_equal(st_r["pos"], sens[5]["root_pos"]) and (ex_r["estimated_root_pos"] == 0).all()
    # estimator and terrain adaptation off: odometry and commanded pitch pass through
    off = pkg.prep_config_default()
    off.use_estimator = 0
    off.use_terrain_adapt = 0
    e.prepare_states(sens[6], off)
    st_f, _, _ = e.get_prepared()
    assert np.array_equal(st_f["pos"], sens[6]["root_pos"]) and np.array_equal(st_f["euler_d"], sens[6]["root_euler_d"])
    e.close()
    e2.close()


@pytest.mark.parametrize("flags", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)])
def test_horizon_extensions_parity(pkg, ob, flags):
    """SURVEY 8f row 4 behind flags: exact discretisation, drifting feet, gait-aware contacts.
    QP matrices and GRFs against the oracle's own restatement of the same extensions."""
    cfg = pkg.config_default()
    cfg.exact_discretization, cfg.foot_drift, cfg.gait_aware = flags
    n = 384
    st = pkg.generate_states(1002, 0, n)
    gait = pkg.generate_gait_inputs(1002, 0, n, 0)
    e = pkg.MpcEngine(cfg, 0)
    e.load_states(st)
    if cfg.gait_aware:
        with pytest.raises(pkg.MpcError) as ei:
            e.build_qp()                                   # gait records are mandatory for such an engine
        assert ei.value.code == pkg.abi.MPC_ERR_STATE
        e.set_gait_inputs(gait)
    e.build_qp()
    worst = 0.0
    for i in range(0, n, 8):
        P, q, l, u = e.get_qp(i)
        Po, qo, lo, uo = ob.mpc_build_qp_ext(cfg, st[i], gait[i])
        worst = max(worst, np.abs(P - Po).max() / np.abs(Po).max(), np.abs(q - qo).max() / np.abs(qo).max())
        assert np.array_equal(l, lo.astype(np.float32)) and np.array_equal(u, uo.astype(np.float32))
    assert worst <= TOL_QP, worst
    e.solve()
    res = e.get_results()
    ref = ob.mpc_compute_grf_ext(cfg, st, gait)
    assert_same_iterates(res, ref, max_flipped=0.0, what=str(flags))
    # the flags do change the answer (otherwise this test checks nothing)
    base = ob.mpc_compute_grf(pkg.config_default(), st)
    assert grf_rel(res["grf"], base["grf"]).max() > 1e-4
    e.close()


@pytest.mark.parametrize("flags", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)])
def test_horizon_extensions_long_horizon(pkg, ob, flags):
    """The same three flags at H = 30 (gen_build_kernel<30> + Riccati solver) against the oracle's restatement."""
    cfg = pkg.config_default()
    cfg.horizon = 30
    cfg.exact_discretization, cfg.foot_drift, cfg.gait_aware = flags
    n = 64
    st = pkg.generate_states(1004, 0, n)
    gait = pkg.generate_gait_inputs(1004, 0, n, 0)
    e = pkg.MpcEngine(cfg, 0)
    e.load_states(st)
    if cfg.gait_aware:
        e.set_gait_inputs(gait)
    e.build_qp()
    worst = 0.0
    for i in (0, 9, 63):
        P, q, l, u = e.get_qp(i)
        Po, qo, lo, uo = ob.mpc_build_qp_ext(cfg, st[i], gait[i])
        worst = max(worst, np.abs(P - Po).max() / np.abs(Po).max(), np.abs(q - qo).max() / np.abs(qo).max())
        assert np.array_equal(l, lo.astype(np.float32)) and np.array_equal(u, uo.astype(np.float32))
    assert worst <= TOL_QP, worst
    e.solve()
    res = e.get_results()
    ref = ob.mpc_compute_grf_ext(cfg, st, gait)
    assert_same_iterates(res, ref, max_flipped=0.0, what=f"H=30 {flags}")
    base_cfg = pkg.config_default()
    base_cfg.horizon = 30
    base = ob.mpc_compute_grf(base_cfg, st)
    assert grf_rel(res["grf"], base["grf"]).max() > 1e-4
    e.close()


def test_run_to_run_determinism(pkg):
    """The solver's warps synchronise through flags, not block barriers: the same batch must give
    bit-identical iteration counts and forces on every run (a slot-recycling race showed up here
    as a few first-problem-per-CTA results changing from run to run)."""
    e = pkg.MpcEngine(pkg.config_default(), 0)
    st = pkg.generate_states(1002, 0, 1024)
    first = e.compute_grf_batch(st).copy()
    for _ in range(40):
        res = e.compute_grf_batch(st)
        assert np.array_equal(res["iters"], first["iters"]) and np.array_equal(res["grf"], first["grf"])
    # fresh engines too (cold instruction cache and untouched shared memory per process are what exposed it)
    for _ in range(5):
        e2 = pkg.MpcEngine(pkg.config_default(), 0)
        res = e2.compute_grf_batch(st)
        assert np.array_equal(res["iters"], first["iters"])
        e2.close()
    e.close()


def test_stream_slots_survive_growth(pkg, ob):
    """Growing the batch keeps the live solvers of the robots that already have a slot; the new
    robots start cold (their first tick is an initSolver)."""
    cfg = pkg.config_hardware()
    n0, n1 = 96, 160
    st = [pkg.generate_stream_states(1006, 0, n1, 44 + t) for t in range(3)]
    e = pkg.MpcEngine(cfg, 0)
    e.stream_step(st[0][:n0])
    e.stream_step(st[1][:n0])
    res = e.stream_step(st[2])                      # 160 robots: 96 warm third ticks + 64 cold first ticks
    warm = ob.mpc_stream(cfg, np.stack([s[:n0] for s in st]))[2]
    cold = ob.mpc_compute_grf(cfg, st[2][n0:])
    assert np.array_equal(res["iters"][:n0], warm["iters"]) and np.array_equal(res["iters"][n0:], cold["iters"])
    assert grf_rel(res["grf"][:n0], warm["grf"]).max() <= TOL_GRF
    assert grf_rel(res["grf"][n0:], cold["grf"]).max() <= TOL_GRF
    e.close()


@pytest.mark.parametrize("H,n,seed", [(10, 1024, 1002), (10, 512, 1003), (30, 96, 1004)])
def test_structured_solver_parity(pkg, ob, H, n, seed):
    """The Riccati-structured solver (riccati_kernel.cuh): the same ADMM with K x = r solved by a
    recursion over the horizon instead of the dense inverse.  Same gates as the dense path."""
    cfg = pkg.config_default()
    cfg.horizon = H
    cfg.structured_solver = 1
    e = pkg.MpcEngine(cfg, 0)
    st = pkg.generate_states(seed, 0, n)
    res = e.compute_grf_batch(st)
    ref = ob.mpc_compute_grf(cfg, st)
    assert (res["status"] == 1).all() and (ref["status"] == 1).all()
    assert grf_rel(res["grf"], ref["grf"]).max() <= TOL_GRF
    assert (res["iters"] == ref["iters"]).mean() >= 0.995
    assert (res["rho_updates"] == ref["rho_updates"]).mean() >= 0.995
    # the full primal solution is feasible like the dense one
    x = e.get_solution(0).astype(np.float64).reshape(-1, 3)
    assert (np.abs(x[:, 0]) <= cfg.mu * x[:, 2] + TOL_CONE * cfg.fz_max).all()
    # the torque map runs behind it too
    tin = pkg.generate_torque_inputs(seed, 0, n)
    e.load_states(st)
    e.set_torque_inputs(tin)
    e.build_qp()
    e.solve()
    _check_torques(ob, st, tin, e.get_results(), e.get_torques())
    e.close()


def test_structured_solver_hardware_weights_and_extensions(pkg, ob):
    cfg = pkg.config_hardware()
    cfg.structured_solver = 1
    cfg.exact_discretization = cfg.foot_drift = cfg.gait_aware = 1
    n = 256
    st = pkg.generate_states(1002, 0, n)
    gait = pkg.generate_gait_inputs(1002, 0, n, 0)
    e = pkg.MpcEngine(cfg, 0)
    e.load_states(st)
    e.set_gait_inputs(gait)
    e.build_qp()
    e.solve()
    res = e.get_results()
    ref = ob.mpc_compute_grf_ext(cfg, st, gait)
    assert_same_iterates(res, ref, max_flipped=0.0, what="riccati + extensions")
    e.close()


@pytest.mark.parametrize("which", ["gazebo", "hardware"])
def test_long_horizon_wrench_engine(pkg, ob, which):
    """H = 30 through wrench_riccati_kernel.cuh (structured_solver = 3, the default at H = 30): the fused build +
    six-input Riccati ADMM.  Same gates as every other engine: oracle iteration counts on every state, GRF inside
    the gate; and the three long-horizon engines (wrench, Riccati on the dense build, dense workspace) agree."""
    cfg = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
    cfg.horizon = 30
    cfg.structured_solver = 3
    n = 160
    st = pkg.generate_states(1004, 0, n)
    e = pkg.MpcEngine(cfg, 0)
    l0 = e.kernel_launches()
    res = e.compute_grf_batch(st).copy()
    assert e.kernel_launches() - l0 == 1                     # ONE kernel from the records to the results
    ref = ob.mpc_compute_grf(cfg, st)
    assert (res["status"] == 1).all()
    assert_same_iterates(res, ref, max_flipped=0.0, what=f"H=30 wrench engine ({which})")
    # the other long-horizon engines on the same states
    for mode, m in ((1, n), (2, 16)):
        c2 = pkg.config_default() if which == "gazebo" else pkg.config_hardware()
        c2.horizon = 30
        c2.structured_solver = mode
        d = pkg.MpcEngine(c2, 0)
        r2 = d.compute_grf_batch(st[:m])
        assert np.array_equal(r2["iters"], res["iters"][:m])
        assert np.abs(r2["grf"] - res["grf"][:m]).max() <= 1e-3
        d.close()
    # full primal solution: feasible, and the first step is what the result record holds
    x = e.get_solution(5).astype(np.float64).reshape(30, 4, 3)
    assert np.maximum(np.abs(x[..., :2]).max(-1) - cfg.mu * x[..., 2], 0).max() <= TOL_CONE * cfg.fz_max
    R = st["rot_mat"][5].astype(np.float64).reshape(3, 3)
    assert np.abs((x[0] @ R).ravel() - res["grf"][5]).max() <= 1e-3 * max(np.abs(res["grf"][5]).max(), 1.0)
    # the fused torque map
    tin = pkg.generate_torque_inputs(1004, 0, n)
    e.load_states(st)
    e.set_torque_inputs(tin)
    e.build_qp()
    e.solve()
    _check_torques(ob, st, tin, e.get_results(), e.get_torques())
    # a repeat is bit-identical; ragged sizes around the resident CTA count
    assert e.compute_grf_batch(st).tobytes() == res.tobytes()
    big = pkg.generate_states(1004, 0, 2 * 148 + 5)
    rb = e.compute_grf_batch(big)
    assert rb[:n].tobytes() == res.tobytes()
    e.close()


def test_long_horizon_wrench_engine_extensions_and_refusals(pkg, ob):
    """foot_drift and gait_aware inside the H = 30 wrench kernel; exact_discretization is the Riccati engine's
    (automatic under structured_solver = 0, refused under 3)."""
    n = 96
    st = pkg.generate_states(1002, 0, n)
    gait = pkg.generate_gait_inputs(1002, 0, n, 0)
    for flags in ((0, 1, 0), (0, 0, 1), (0, 1, 1)):
        cfg = pkg.config_default()
        cfg.horizon = 30
        cfg.structured_solver = 3
        cfg.exact_discretization, cfg.foot_drift, cfg.gait_aware = flags
        e = pkg.MpcEngine(cfg, 0)
        e.load_states(st)
        if cfg.gait_aware:
            e.set_gait_inputs(gait)
        e.build_qp()
        e.solve()
        ref = ob.mpc_compute_grf_ext(cfg, st, gait)
        assert_same_iterates(e.get_results(), ref, max_flipped=0.0, what=f"H=30 wrench engine, flags {flags}")
        e.close()
    cfg = pkg.config_default()
    cfg.horizon = 30
    cfg.exact_discretization = 1
    cfg.structured_solver = 3
    with pytest.raises(pkg.MpcError):
        pkg.MpcEngine(cfg, 0)
    cfg.structured_solver = 0                                 # auto: falls back to the Riccati engine
    e = pkg.MpcEngine(cfg, 0)
    res = e.compute_grf_batch(st[:32])
    ref = ob.mpc_compute_grf_ext(cfg, st[:32], gait[:32])
    assert_same_iterates(res, ref, max_flipped=0.0, what="H=30 auto with exact_discretization")
    e.close()


@pytest.mark.parametrize("H", [10, 30])
def test_solver_settings_sweep(pkg, ob, H):
    """The fused kernels follow the oracle under every OSQP setting the ABI exposes, not only the defaults: no
    equilibration, fixed rho, loose tolerance, iteration limits that end mid-way (status 2 / -2), no relaxation,
    other check / adaptation intervals, other initial rho, fewer Ruiz passes."""
    cases = [dict(scaling=0), dict(adaptive_rho=0), dict(eps_abs=1e-3, eps_rel=1e-3), dict(max_iter=60),
             dict(max_iter=110), dict(alpha=1.0), dict(check_termination=10, adaptive_rho_interval=20), dict(rho=1.0),
             dict(scaling=3)]
    st = pkg.generate_states(1004, 0, 64)
    seen = set()
    for kw in cases:
        cfg = pkg.config_hardware()
        cfg.horizon = H
        for k, v in kw.items():
            setattr(cfg.osqp, k, v)
        e = pkg.MpcEngine(cfg, 0)
        r = e.compute_grf_batch(st)
        e.close()
        ref = ob.mpc_compute_grf(cfg, st)
        assert_same_iterates(r, ref, max_flipped=0.0, what=f"H={H} {kw}")
        seen |= set(r["status"].tolist())
    assert {1, 2, -2} <= seen          # the iteration limits really ended solves mid-way


@pytest.mark.parametrize("H", [10, 30])
def test_degenerate_foot_geometry(pkg, ob, H):
    """All four feet at ONE point: the wrench map B6c has rank 3, the 6 x 6 matrices N_k of the wrench-space
    kernels are singular (zero Cholesky pivots).  The fused kernels still produce the oracle's iterates."""
    cfg = pkg.config_default()
    cfg.horizon = H
    st = pkg.generate_states(1004, 0, 32)
    fp = st["foot_pos_abs"].reshape(-1, 4, 3).copy()
    fp[:, :, :] = fp[:, :1, :]
    st["foot_pos_abs"] = fp.reshape(st["foot_pos_abs"].shape)
    e = pkg.MpcEngine(cfg, 0)
    r = e.compute_grf_batch(st)
    e.close()
    assert_same_iterates(r, ob.mpc_compute_grf(cfg, st), max_flipped=0.0, what=f"degenerate feet, H={H}")


def test_long_horizon_dense_workspace_path(pkg, ob):
    """H = 30 through the dense K^-1-in-L2-workspace kernels (structured_solver = 2), kept as the
    independent second implementation of the long horizon."""
    cfg = pkg.config_default()
    cfg.horizon = 30
    cfg.structured_solver = 2
    e = pkg.MpcEngine(cfg, 0)
    st = pkg.generate_states(1004, 0, 24)
    res = e.compute_grf_batch(st)
    ref = ob.mpc_compute_grf(cfg, st)
    assert (res["status"] == 1).all()
    assert np.array_equal(res["iters"], ref["iters"]) and grf_rel(res["grf"], ref["grf"]).max() <= TOL_GRF
    e.close()


def test_two_engines_two_host_threads(pkg):
    """The double-buffered use INTEGRATION.md recommends: two engines, one host thread each, batches
    alternating between them while the kernels of both share the SMs.  Results must be the bits of a
    single engine working through the same batches alone."""
    import threading
    cfg = pkg.config_default()
    batches = [pkg.generate_states(1002, 700 * b, 700) for b in range(6)]
    solo = pkg.MpcEngine(cfg, 0)
    want = [solo.compute_grf_batch(b).copy() for b in batches]
    solo.close()
    engs = [pkg.MpcEngine(cfg, 0), pkg.MpcEngine(cfg, 0)]
    got = [None] * len(batches)
    errs = []

    def worker(i):
        try:
            for b in range(i, len(batches), 2):
                got[b] = engs[i].compute_grf_batch(batches[b]).copy()
        except Exception as ex:  # surfaced below: an exception in a thread must fail the test
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in engs:
        e.close()
    assert not errs, errs
    for b in range(len(batches)):
        assert got[b].tobytes() == want[b].tobytes()


def test_config4_full_size_h30(pkg, ob):
    """BASELINE configs[3] at its full size: 8192 states, H = 30.  Every state: solved, plausible iteration
    count, feasible first-step force; a 1/64 sample against the oracle; a repeat is bit-identical."""
    cfg = pkg.config_default()
    cfg.horizon = 30
    e = pkg.MpcEngine(cfg, 0)
    st = pkg.generate_states(1004, 0, 8192)
    res = e.compute_grf_batch(st).copy()
    assert (res["status"] == 1).all()
    assert res["iters"].min() >= 50 and res["iters"].max() <= 2000 and (res["iters"] % 25 == 0).all()
    Rm = st["rot_mat"].reshape(-1, 3, 3).astype(np.float64)
    fw = np.einsum("nij,nlj->nli", Rm, res["grf"].reshape(-1, 4, 3).astype(np.float64))
    cmask = st["contacts"].astype(np.float64)
    assert (fw[..., 2] >= -TOL_CONE * 180).all() and (fw[..., 2] <= 180.0 * cmask + TOL_CONE * 180 + 2e-3).all()
    assert (np.abs(fw[..., :2]).max(-1) <= 0.3 * fw[..., 2] + TOL_CONE * 180 + 2e-3).all()
    idx = np.arange(0, 8192, 64)
    ref = ob.mpc_compute_grf(cfg, st[idx])
    assert_same_iterates(res[idx], ref, max_flipped=0.0, what="config 4 sample")
    assert e.compute_grf_batch(st).tobytes() == res.tobytes()
    e.close()


def test_wrench_kernels_agree(pkg, ob, tmp_path):
    """The two register layouts of the H = 10 wrench-space kernel -- 4 x 8 tiles (default) and half rows
    (MPC_WRENCH_TILE=0, read once per process, hence the second process) -- produce the same iteration counts and
    rho updates on every state, cold and warm-started, and forces equal far inside the gate."""
    import subprocess
    import sys
    n = 2048
    cfg = pkg.config_default()
    st = pkg.generate_states(1002, 300, n)
    w = [pkg.generate_stream_states(1006, 0, 256, 44 + t) for t in range(3)]
    e = pkg.MpcEngine(cfg, 0)
    res = e.compute_grf_batch(st).copy()
    warm = [e.stream_step(s).copy() for s in w]
    e.close()
    out = tmp_path / "half.npz"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import go1_qp_mpc_controller_b200 as pkg;"
        "e = pkg.MpcEngine(pkg.config_default(), 0);"
        "r = e.compute_grf_batch(pkg.generate_states(1002, 300, %d)).copy();"
        "w = [e.stream_step(pkg.generate_stream_states(1006, 0, 256, 44 + t)).copy() for t in range(3)];"
        "np.savez(%r, r=r, w0=w[0], w1=w[1], w2=w[2]); e.close()" % (root, n, str(out)))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, MPC_WRENCH_TILE="0"), timeout=600)
    ref = np.load(out)
    for a, b in [(res, ref["r"]), (warm[0], ref["w0"]), (warm[1], ref["w1"]), (warm[2], ref["w2"])]:
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"])
        assert np.array_equal(a["rho_updates"], b["rho_updates"])
        assert grf_rel(a["grf"], b["grf"].astype(np.float64)).max() <= 1e-4


def test_balance_kernels_agree(pkg, ob, tmp_path):
    """The two layouts of the stance-balance QP -- four lanes per problem (default) and one warp per problem
    (MPC_BALANCE_KERNEL=warp, read once per process, hence the second process) -- carry the same arithmetic:
    identical iteration counts and rho updates, forces equal to fp32 rounding of the last bits."""
    import subprocess
    import sys
    n = 4096
    bcfg = pkg.balance_config_default()
    st = pkg.generate_balance_states(1005, 77, n)
    be = pkg.MpcEngine(bcfg, 0, balance=True)
    res = be.compute_grf_batch(st).copy()
    be.close()
    out = tmp_path / "warp.npy"
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import go1_qp_mpc_controller_b200 as pkg;"
        "e = pkg.MpcEngine(pkg.balance_config_default(), 0, balance=True);"
        "np.save(%r, e.compute_grf_batch(pkg.generate_balance_states(1005, 77, %d))); e.close()"
        % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(out), n))
    env = dict(os.environ, MPC_BALANCE_KERNEL="warp")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
    ref = np.load(out)
    assert np.array_equal(res["status"], ref["status"])
    assert np.array_equal(res["iters"], ref["iters"]) and np.array_equal(res["rho_updates"], ref["rho_updates"])
    assert grf_rel(res["grf"], ref["grf"].astype(np.float64)).max() <= 1e-5
    # ragged sizes: not a multiple of the eight problems of a warp, fewer than one warp
    be = pkg.MpcEngine(bcfg, 0, balance=True)
    for m in (1, 7, 9, 33, 1001):
        r = be.compute_grf_batch(st[:m])
        assert r.tobytes() == res[:m].tobytes(), m
    be.close()
    # a bad record (NaN) is reported unsolved with zero force (A1RobotControl.cpp:441-443) and its seven
    # warp-mates do not notice
    be = pkg.MpcEngine(bcfg, 0, balance=True)
    bad = st[:64].copy()
    bad["pos"][5] = np.nan
    r = be.compute_grf_batch(bad)
    assert r["status"][5] != 1 and (r["grf"][5] == 0).all()
    ok = np.arange(64) != 5
    assert r[ok].tobytes() == res[:64][ok].tobytes()
    be.close()


def test_config5_full_size_balance(pkg, ob):
    """BASELINE configs[4] at its full size: 1 000 000 stance-balance QPs (on one GPU here; bench.py times
    the per-GPU share).  Status / iteration histogram on all of them, the oracle on a 1/1024 sample."""
    bcfg = pkg.balance_config_default()
    be = pkg.MpcEngine(bcfg, 0, balance=True)
    n = 1_000_000
    st = pkg.generate_balance_states(1005, 0, n)
    res = be.compute_grf_batch(st)
    codes, counts = np.unique(res["status"], return_counts=True)
    assert set(codes.tolist()) <= {1, 2, -2} and counts[codes == 1][0] >= 0.999 * n, dict(zip(codes.tolist(), counts.tolist()))
    assert (res["iters"] % 25 == 0).all() and res["iters"].min() >= 25 and res["iters"].max() <= bcfg.osqp.max_iter
    assert 100 <= res["iters"].mean() <= 400
    idx = np.arange(0, n, 1024)
    ref = ob.balance_compute_grf(bcfg, st[idx])
    assert_same_iterates(res[idx], ref, max_flipped=0.005, what="config 5 sample")
    # contact-gated: a swing leg carries no force in the body frame
    sw = st["contacts"] == 0
    assert np.abs(res["grf"].reshape(-1, 4, 3)[sw]).max() < 2e-2
    be.close()


def test_mirror_compute_grf_keeps_one_warm_solver(pkg, ob):
    """The Python mirror's A1RobotControl.compute_grf: one persistent solver (A1RobotControl.h:67), 20
    ticks of one robot across a trot swap against the oracle's MpcStream; a weight change in the state
    is one more Hessian update of the SAME solver (A1RobotControl.cpp:447 re-reads the weights)."""
    cfg = pkg.config_default()
    T, robot = 20, 3
    st = np.stack([pkg.generate_stream_states(1006, robot, 1, 38 + t) for t in range(T)])
    ref = ob.mpc_stream(cfg, st)
    ctl = pkg.A1RobotControl(pkg.config_default())
    s = pkg.A1CtrlStates()
    s.robot_mass = cfg.mass
    s.q_weights = np.array(cfg.q_weights[:])
    s.r_weights = np.array(cfg.r_weights[:])
    s.a1_trunk_inertia = np.array(cfg.inertia[:]).reshape(3, 3)
    iters = []
    for t in range(T):
        rec = st[t, 0]
        s.root_euler, s.root_pos = rec["euler"].astype(float), rec["pos"].astype(float)
        s.root_ang_vel, s.root_lin_vel = rec["ang_vel"].astype(float), rec["lin_vel"].astype(float)
        s.root_euler_d = rec["euler_d"].astype(float)
        s.root_pos_d = np.array([0.0, 0.0, float(rec["pos_d_z"])])
        s.root_lin_vel_d, s.root_ang_vel_d = rec["lin_vel_d"].astype(float), rec["ang_vel_d"].astype(float)
        s.root_rot_mat = rec["rot_mat"].astype(float).reshape(3, 3)
        s.foot_pos_abs = rec["foot_pos_abs"].astype(float).reshape(4, 3).T
        s.contacts = [bool(c) for c in rec["contacts"]]
        grf = ctl.compute_grf(s, 0.0025)
        assert ctl.last_status == 1 and ctl.last_iters == ref["iters"][t, 0], (t, ctl.last_iters, ref["iters"][t, 0])
        want = ref["grf"][t, 0].reshape(4, 3).T
        assert np.linalg.norm(grf - want) / max(np.linalg.norm(want), 1.0) <= TOL_GRF, t
        iters.append(ctl.last_iters)
    assert np.mean(iters[1:9]) < 0.7 * iters[0]
    # new weights: same engine, same live solver (no cold restart)
    eng_before = ctl._engine
    s.q_weights = s.q_weights * 1.5
    ctl.compute_grf(s, 0.0025)
    assert ctl._engine is eng_before and ctl.last_iters < iters[0]
    ctl.reset_solver()
    ctl.compute_grf(s, 0.0025)
    assert ctl.last_iters >= 100                                   # cold again


@pytest.mark.parametrize("H", [10, 30])
def test_warm_slot_fault_containment_and_slot_reset(pkg, ob, H):
    """One bad record (NaN) poisons only its own robot slot, and only for that tick: the kernel leaves the
    slot dead, so the robot's next tick is an initSolver; every other robot keeps its warm solver.
    mpc_stream_reset_slots forgets chosen robots only.  Both fused kernels (H = 10 and H = 30)."""
    cfg = pkg.config_hardware()
    cfg.horizon = H
    N = 96 if H == 10 else 40
    st = [pkg.generate_stream_states(1006, 0, N, 44 + t) for t in range(4)]
    ref = ob.mpc_stream(cfg, np.stack(st))
    e = pkg.MpcEngine(cfg, 0)
    e.stream_step(st[0])
    bad = st[1].copy()
    bad["pos"][5] = np.nan
    r1 = e.stream_step(bad)
    assert (r1["grf"][5] == 0).all() and r1["status"][5] != 1      # NaN guard: zero force, not solved
    ok = np.arange(N) != 5
    assert np.array_equal(r1["iters"][ok], ref["iters"][1][ok])
    r2 = e.stream_step(st[2])
    cold2 = ob.mpc_compute_grf(cfg, st[2][5:6])
    assert r2["status"][5] == 1 and r2["iters"][5] == cold2["iters"][0]          # recovered by a cold start
    assert np.array_equal(r2["iters"][ok], ref["iters"][2][ok])                   # nobody else noticed
    # per-slot reset
    e.stream_reset_slots([7, 9])
    r3 = e.stream_step(st[3])
    cold3 = ob.mpc_compute_grf(cfg, st[3][[7, 9]])
    assert np.array_equal(r3["iters"][[7, 9]], cold3["iters"])
    keep = ok & (np.arange(N) != 7) & (np.arange(N) != 9)
    assert np.array_equal(r3["iters"][keep], ref["iters"][3][keep])
    e.close()
