"""CPU tests of the oracle (the checker) -- no GPU needed.

The reference holds no golden vectors for this path (parity unpinned), so
the oracle is pinned by what CAN be checked here:
  * committed golden vectors (regression of the oracle itself),
  * the survey's independent numpy probe of test/test_mpc.cpp's input,
  * structural identities of the QP build,
  * an independent numpy restatement of ConvexMpc.cpp,
  * a KKT optimality certificate of the tight-tolerance solution (P is SPD, so a
    KKT point is THE optimum whichever algorithm found it).
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg_test_mpc(pkg):
    cfg = pkg.config_default()
    cfg.mass = 15.0
    for i, v in enumerate([0.0158533, 0, 0, 0, 0.0377999, 0, 0, 0, 0.0456542]):
        cfg.inertia[i] = v
    for i, v in enumerate([1.0, 1.0, 1.0, 0.0, 0.0, 50.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0]):
        cfg.q_weights[i] = v
    for i in range(12):
        cfg.r_weights[i] = 1e-6
    return cfg


# ---------------------------------------------------------------------------
# independent numpy restatement of ConvexMpc.cpp + A1RobotControl.cpp:452-518
# ---------------------------------------------------------------------------
def numpy_build(cfg, rec):
    H = cfg.horizon
    q_w = np.array(cfg.q_weights[:])
    r_w = np.array(cfg.r_weights[:])
    Q = np.diag(np.tile(2 * q_w, H))
    R = np.diag(np.tile(2 * r_w, H))
    euler = rec["euler"].astype(np.float64)
    Rm = rec["rot_mat"].astype(np.float64).reshape(3, 3)
    cy, sy = np.cos(euler[2]), np.sin(euler[2])
    Ac = np.zeros((13, 13))
    Ac[0:3, 6:9] = [[cy, sy, 0], [-sy, cy, 0], [0, 0, 1]]
    Ac[3:6, 9:12] = np.eye(3)
    Ac[11, 12] = 1
    I = np.array(cfg.inertia[:]).reshape(3, 3)
    Iw_inv = np.linalg.inv(Rm @ I @ Rm.T)
    foot = rec["foot_pos_abs"].astype(np.float64).reshape(4, 3)
    Bc = np.zeros((13, 12))
    for i in range(4):
        v = foot[i]
        sk = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        Bc[6:9, 3 * i:3 * i + 3] = Iw_inv @ sk
        Bc[9:12, 3 * i:3 * i + 3] = np.eye(3) / cfg.mass
    Ad = np.eye(13) + Ac * cfg.dt
    Bd = Bc * cfg.dt
    Aqp = np.vstack([np.linalg.matrix_power(Ad, i + 1) for i in range(H)])
    Bqp = np.zeros((13 * H, 12 * H))
    for i in range(H):
        for j in range(i + 1):
            Bqp[13 * i:13 * i + 13, 12 * j:12 * j + 12] = np.linalg.matrix_power(Ad, i - j) @ Bd
    x0 = np.concatenate([euler, rec["pos"], rec["ang_vel"], rec["lin_vel"], [-9.8]]).astype(np.float64)
    vdw = Rm @ rec["lin_vel_d"].astype(np.float64)
    xref = np.zeros(13 * H)
    for i in range(H):
        xref[13 * i:13 * i + 13] = [rec["euler_d"][0], rec["euler_d"][1],
                                    euler[2] + float(rec["ang_vel_d"][2]) * cfg.dt * (i + 1),
                                    float(rec["pos"][0]) + vdw[0] * cfg.dt * (i + 1),
                                    float(rec["pos"][1]) + vdw[1] * cfg.dt * (i + 1), rec["pos_d_z"],
                                    rec["ang_vel_d"][0], rec["ang_vel_d"][1], rec["ang_vel_d"][2],
                                    vdw[0], vdw[1], 0.0, -9.8]
    P = Bqp.T @ Q @ Bqp + R
    g = Bqp.T @ Q @ (Aqp @ x0 - xref)
    return dict(Ac=Ac, Ad=Ad, Bd=Bd, Aqp=Aqp, Bqp=Bqp, P=P, q=g)


def test_golden_regression(pkg, ob):
    for name, cfg in (("gazebo", pkg.config_default()), ("hardware", pkg.config_hardware())):
        g = np.load(os.path.join(GOLD, f"mpc_{name}.npz"))
        states = g["states"]
        res, sol = ob.mpc_compute_grf(cfg, states, want_solutions=True)
        assert np.array_equal(res["iters"], g["iters"])
        assert np.array_equal(res["status"], g["status"])
        assert np.array_equal(res["rho_updates"], g["rho_updates"])
        np.testing.assert_allclose(res["grf"], g["grf"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(sol, g["solutions"], rtol=0, atol=1e-9)
        P, q, l, u = ob.mpc_build_qp(cfg, states[0])
        np.testing.assert_allclose(P, g["P0"], rtol=1e-13, atol=1e-18)
        np.testing.assert_allclose(q, g["q0"], rtol=1e-12, atol=1e-16)
        assert np.array_equal(l, g["l0"]) and np.array_equal(u, g["u0"])


def test_golden_regression_h30(pkg, ob):
    g = np.load(os.path.join(GOLD, "mpc_h30.npz"))
    cfg = pkg.config_default()
    cfg.horizon = 30
    res = ob.mpc_compute_grf(cfg, g["states"])
    assert np.array_equal(res["iters"], g["iters"]) and np.array_equal(res["status"], g["status"])
    np.testing.assert_allclose(res["grf"], g["grf"], rtol=0, atol=1e-9)
    P, q, l, u = ob.mpc_build_qp(cfg, g["states"][0])
    assert P.shape == (360, 360) and abs(np.linalg.eigvalsh(P)[0] - 2e-7) < 1e-11


def test_generator_known_answer(pkg):
    g = np.load(os.path.join(GOLD, "generator.npz"))
    assert pkg.generate_states(1001, 0, 4).tobytes() == g["s1001"].tobytes()
    assert pkg.generate_states(1003, 65530, 4).tobytes() == g["s1003_off"].tobytes()
    assert pkg.generate_balance_states(1005, 999_998, 2).tobytes() == g["b1005"].tobytes()
    # counter based: any shard equals the same slice of a bigger batch
    a = pkg.generate_states(1003, 0, 64)
    b = pkg.generate_states(1003, 40, 24)
    assert a[40:].tobytes() == b.tobytes()
    # contact mix: trot pairs or four-stance only, rotation matrices orthonormal
    s = pkg.generate_states(1002, 0, 2000)
    c = s["contacts"]
    assert np.all((c[:, 0] == c[:, 3]) & (c[:, 1] == c[:, 2]) & (c.sum(1) >= 2))
    frac4 = (c.sum(1) == 4).mean()
    assert 0.07 < frac4 < 0.13
    Rm = s["rot_mat"].reshape(-1, 3, 3).astype(np.float64)
    assert np.abs(Rm @ Rm.transpose(0, 2, 1) - np.eye(3)).max() < 1e-6


def test_survey_probe_known_answer(pkg, ob):
    """SURVEY.md 8c [PROBE]: an independent numpy restatement of test/test_mpc.cpp's input."""
    g = np.load(os.path.join(GOLD, "test_mpc_case.npz"))
    cfg = _cfg_test_mpc(pkg)
    P, q, l, u = ob.mpc_build_qp(cfg, g["state"][0])
    assert abs(P[0, 0] - 1.354e-3) < 1e-6 and abs(P[0, 1] + 1.530e-3) < 1e-6
    assert abs(np.trace(P) - 6.8275e-2) < 1e-6 and abs(q[2] + 4.52e-4) < 1e-6
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-5
    r = ob.mpc_compute_grf(cfg, g["state"])
    assert r["iters"][0] == 50
    np.testing.assert_allclose(r["grf"][0][[1, 2, 7, 8]], [-12.782, 42.606, -12.782, 42.606], atol=2e-3)
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-10
    cfg.osqp.max_iter = 20000
    r, sol = ob.mpc_compute_grf(cfg, g["state"], want_solutions=True)
    np.testing.assert_allclose(r["grf"][0][[1, 2, 7, 8]], [-12.837, 42.790, -12.837, 42.790], atol=1e-3)
    assert np.abs(r["grf"][0][[3, 4, 5, 9, 10, 11]]).max() < 1e-6  # swing legs FR, RR carry nothing
    x = sol[0]
    assert abs(0.5 * x @ P @ x + q @ x + 0.1364844) < 1e-6          # objective, SURVEY.md 8c


def test_structural_identities(pkg, ob):
    cfg = pkg.config_default()
    states = pkg.generate_states(1002, 100, 6)
    for rec in states:
        im = ob.mpc_build_intermediates(cfg, rec)
        Ad, Bd, Aqp, Bqp = im["A_d"], im["B_d"], im["A_qp"], im["B_qp"]
        Ac = (Ad - np.eye(13)) / cfg.dt
        assert np.abs(np.linalg.matrix_power(Ac, 3)).max() == 0.0          # A_c^3 = 0
        dt = cfg.dt
        for k in range(1, 11):                                               # closed form of A_d^k
            ref = np.eye(13) + k * dt * Ac + k * (k - 1) / 2 * dt * dt * (Ac @ Ac)
            np.testing.assert_allclose(Aqp[13 * (k - 1):13 * k], ref, rtol=0, atol=1e-15)
        for i in range(10):                                                  # block Toeplitz B_qp
            for j in range(10):
                blk = Bqp[13 * i:13 * i + 13, 12 * j:12 * j + 12]
                if j > i:
                    assert np.abs(blk).max() == 0.0
                else:
                    np.testing.assert_allclose(blk, Bd + (i - j) * dt * (Ac @ Bd), rtol=0, atol=1e-15)
        P, q, l, u = ob.mpc_build_qp(cfg, rec)
        assert np.abs(P - P.T).max() < 1e-17
        ev = np.linalg.eigvalsh(P)
        assert abs(ev[0] - 2 * min(cfg.r_weights[:])) < 1e-12                # lambda_min = 2 min(r)
        # bounds: pyramid rows one-sided, fz in [0, 180 c], contacts replicated over the horizon
        c = rec["contacts"]
        for h in range(10):
            for leg in range(4):
                b = 20 * h + 5 * leg
                assert list(l[b:b + 4]) == [0.0, -1e30, 0.0, -1e30] and list(u[b:b + 4]) == [1e30, 0.0, 1e30, 0.0]
                assert l[b + 4] == 0.0 and u[b + 4] == 180.0 * c[leg]


def test_oracle_matches_numpy_restatement(pkg, ob):
    for cfg in (pkg.config_default(), pkg.config_hardware()):
        states = pkg.generate_states(1002, 7, 5)
        for rec in states:
            nb = numpy_build(cfg, rec)
            im = ob.mpc_build_intermediates(cfg, rec)
            P, q, l, u = ob.mpc_build_qp(cfg, rec)
            np.testing.assert_allclose(im["A_d"], nb["Ad"], rtol=0, atol=1e-15)
            np.testing.assert_allclose(im["B_d"], nb["Bd"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(im["B_qp"], nb["Bqp"], rtol=1e-11, atol=1e-15)
            assert np.abs(P - nb["P"]).max() / np.abs(P).max() < 1e-12
            assert np.abs(q - nb["q"]).max() / np.abs(q).max() < 1e-11


def _kkt_violation(P, q, A, l, u, x, y):
    """Stationarity, primal feasibility, dual sign and complementarity of (x, y)."""
    stat = np.abs(P @ x + q + A.T @ y).max()
    Ax = A @ x
    prim = max(np.maximum(l - Ax, 0).max(), np.maximum(Ax - u, 0).max())
    fin_l, fin_u = l > -1e29, u < 1e29
    # y_i > 0 only at an active upper bound, y_i < 0 only at an active lower bound
    comp_u = (np.maximum(y, 0) * np.where(fin_u, u - Ax, 0.0)).max()
    comp_l = (np.maximum(-y, 0) * np.where(fin_l, Ax - l, 0.0)).max()
    comp = max(abs(comp_u), abs(comp_l))
    # a multiplier pushing against an infinite bound is never allowed
    sign = max(np.maximum(y, 0)[~fin_u].max(initial=0.0), np.maximum(-y, 0)[~fin_l].max(initial=0.0))
    return stat, prim, comp, sign


def test_kkt_certificate_of_tight_solution(pkg, ob):
    cfg = pkg.config_default()
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-10
    cfg.osqp.max_iter = 40000
    A = ob.constraint_matrix(10, cfg.mu)
    states = pkg.generate_states(1002, 0, 12)
    for rec in states:
        P, q, l, u = ob.mpc_build_qp(cfg, rec)
        x, y, info = ob.osqp_solve_mpc(cfg, P, q, l, u)
        assert info["status"] == 1
        stat, prim, comp, sign = _kkt_violation(P, q, A, l, u, x, y)
        scale = max(np.abs(q).max(), np.abs(P @ x).max())
        assert stat < 1e-8 * max(scale, 1.0)
        assert prim < 1e-6 and sign < 1e-9
        assert comp < 1e-6 * max(np.abs(y).max(), 1.0) * 180.0


def test_eps_1e5_is_close_to_optimum_for_trot(pkg, ob):
    """trot states land within ~1e-4 of the optimum at eps 1e-5, four-stance do not."""
    cfg = pkg.config_default()
    states = pkg.generate_states(1002, 0, 48)
    r5 = ob.mpc_compute_grf(cfg, states)
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-10
    cfg.osqp.max_iter = 40000
    rt = ob.mpc_compute_grf(cfg, states)
    den = np.maximum(np.linalg.norm(rt["grf"], axis=1), 1.0)
    rel = np.linalg.norm(r5["grf"] - rt["grf"], axis=1) / den
    trot = states["contacts"].sum(1) == 2
    assert np.median(rel[trot]) < 2e-4
    assert (r5["status"] == 1).all() and (rt["status"] == 1).all()
    assert r5["iters"].min() >= 75 and r5["iters"].max() <= 600   # 150-450 typical


def test_balance_oracle(pkg, ob):
    g = np.load(os.path.join(GOLD, "balance.npz"))
    bcfg = pkg.balance_config_default()
    res = ob.balance_compute_grf(bcfg, g["states"])
    assert np.array_equal(res["iters"], g["iters"]) and np.array_equal(res["status"], g["status"])
    np.testing.assert_allclose(res["grf"], g["grf"], rtol=0, atol=1e-9)
    P, q, l, u = ob.balance_build_qp(bcfg, g["states"][0])
    np.testing.assert_allclose(P, g["P0"], rtol=1e-13)
    assert np.abs(P - P.T).max() < 1e-12 and np.linalg.eigvalsh(P)[0] >= bcfg.R - 1e-12
    # rows 0-3 select fz with contact-gated bounds, rows 4-19 are one-sided (A1RobotControl.cpp:28-48)
    c = g["states"][0]["contacts"]
    assert list(l[:4]) == [0.0] * 4 and list(u[:4]) == list(180.0 * c)
    assert np.all(l[4:] == -1e30) and np.all(u[4:] == 0.0)
    # solved forces respect the friction pyramid mu = 0.7 in the world frame (rotate back with R)
    Rm = g["states"]["rot_mat"].reshape(-1, 3, 3).astype(np.float64)
    f_body = res["grf"].reshape(-1, 4, 3)
    f_world = np.einsum("nij,nlj->nli", Rm, f_body)
    viol = np.maximum(np.abs(f_world[..., :2]).max(-1) - 0.7 * f_world[..., 2], 0).max()
    assert viol < 2e-2  # ADMM x-iterate satisfies constraints to the primal residual only
    assert f_world[..., 2].min() > -1e-2 and f_world[..., 2].max() < 180.0 + 1e-2


def test_f32_model_drifts(pkg, ob):
    """Documents WHY the device path computes in f64: the same algorithm in f32 loses the gate."""
    cfg = pkg.config_default()
    states = pkg.generate_states(1002, 0, 256)
    r64 = ob.mpc_compute_grf(cfg, states)
    r32 = ob.mpc_compute_grf(cfg, states, f32=True)
    den = np.maximum(np.linalg.norm(r64["grf"], axis=1), 1.0)
    rel = np.linalg.norm(r32["grf"] - r64["grf"], axis=1) / den
    assert (rel > 1e-3).mean() > 0.002       # f32 misses the 1e-3 gate on a visible fraction
    assert (r32["iters"] == r64["iters"]).mean() < 0.99


# ---- warm-started streaming (A1RobotControl.cpp:522-538) --------------------------------

def test_stream_generator_tick0_is_the_draw(pkg):
    assert pkg.generate_stream_states(1002, 7, 33, 0).tobytes() == pkg.generate_states(1002, 7, 33).tobytes()
    a = pkg.generate_stream_states(1002, 0, 64, 48)
    b = pkg.generate_states(1002, 0, 64)
    trot = b["contacts"].sum(1) == 2
    assert (a["contacts"][trot] == 1 - b["contacts"][trot]).all()       # pairs swapped after 48 ticks
    assert (a["contacts"][~trot] == b["contacts"][~trot]).all()
    R = a["rot_mat"].reshape(-1, 3, 3).astype(np.float64)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-6


@pytest.mark.parametrize("name", ["gazebo", "hardware"])
def test_stream_golden_regression(pkg, ob, name):
    g = np.load(os.path.join(GOLD, f"stream_{name}.npz"))
    cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
    assert g["states"].tobytes() == np.stack(
        [pkg.generate_stream_states(1006, 0, 48, 44 + t) for t in range(8)]).tobytes()
    res = ob.mpc_stream(cfg, g["states"])
    assert np.array_equal(res["iters"], g["iters"]) and np.array_equal(res["status"], g["status"])
    assert np.abs(res["grf"] - g["grf"]).max() <= 1e-9 * np.abs(g["grf"]).max()


def test_stream_semantics(pkg, ob):
    """First tick = initSolver = the cold path; later ticks start from the previous iterates and
    cost fewer iterations; at a tight tolerance warm and cold ticks reach the same optimum."""
    cfg = pkg.config_hardware()
    N, T = 24, 4
    st = np.stack([pkg.generate_stream_states(1002, 0, N, t) for t in range(T)])
    warm = ob.mpc_stream(cfg, st)
    cold0 = ob.mpc_compute_grf(cfg, st[0])
    assert np.array_equal(warm["iters"][0], cold0["iters"]) and np.array_equal(warm["grf"][0], cold0["grf"])
    cold = np.stack([ob.mpc_compute_grf(cfg, st[t]) for t in range(T)])
    assert (warm["status"] == 1).all()
    assert warm["iters"][1:].mean() < 0.7 * cold["iters"][1:].mean()
    cfg.osqp.eps_abs = cfg.osqp.eps_rel = 1e-9
    cfg.osqp.max_iter = 50000
    warm = ob.mpc_stream(cfg, st)
    cold = np.stack([ob.mpc_compute_grf(cfg, st[t]) for t in range(T)])
    assert np.abs(warm["grf"] - cold["grf"]).max() <= 1e-5 * np.abs(cold["grf"]).max()


# ---- torque map (compute_joint_torques, A1RobotControl.cpp:289-319) -----------------------

def test_torque_map_oracle(pkg, ob):
    n = 64
    st = pkg.generate_states(1002, 0, n)
    tin = pkg.generate_torque_inputs(1002, 0, n)
    rng = np.random.default_rng(7)
    grf = rng.uniform(-50, 150, size=(n, 12))
    tau, mask = ob.torque_map(st, tin, grf)
    assert (mask == 0).all()
    for i in range(n):
        for leg in range(4):
            J = tin["j_foot"][i][9 * leg:9 * leg + 9].astype(np.float64).reshape(3, 3)
            g = tin["torques_gravity"][i][3 * leg:3 * leg + 3].astype(np.float64)
            t = tau[i, 3 * leg:3 * leg + 3] - g
            if st["contacts"][i][leg] != 0:
                assert np.allclose(t, -J.T @ grf[i, 3 * leg:3 * leg + 3], rtol=1e-13, atol=1e-13)
            else:
                rhs = tin["km_foot"][i].astype(np.float64) * tin["foot_forces_kin"][i][3 * leg:3 * leg + 3].astype(np.float64)
                assert np.allclose(J @ t, rhs, rtol=1e-10, atol=1e-10)       # J tau = km .* f_kin
                assert np.allclose(t, np.linalg.solve(J, rhs), rtol=1e-9, atol=1e-12)
    # singular Jacobian on a swing leg: NaN components are flagged, not written
    swing = np.argwhere(st["contacts"] == 0)
    i, leg = swing[0]
    tin2 = tin.copy()
    tin2["j_foot"][i][9 * leg:9 * leg + 9] = 0.0
    tau2, mask2 = ob.torque_map(st, tin2, grf)
    assert (mask2[i] >> (3 * leg)) & 7 == 7 and (np.delete(mask2, i) == 0).all()


def test_torque_input_generator(pkg):
    t = pkg.generate_torque_inputs(1002, 5, 16)
    assert t.tobytes() == pkg.generate_torque_inputs(1002, 0, 32)[5:21].tobytes()   # counter based
    J = t["j_foot"].reshape(16, 4, 3, 3).astype(np.float64)
    assert (J[:, :, 0, 0] == 0).all()                       # the hip joint does not move the foot along x
    assert (np.abs(np.linalg.det(J)) > 1e-3).all()          # drawn away from the knee singularity
    assert np.allclose(t["torques_gravity"][0], [0.8, 0, 0, -0.8, 0, 0, 0.8, 0, 0, -0.8, 0, 0])


# ---- upstream state preparation (GazeboA1ROS.cpp:262-288, A1BasicEKF.cpp, terrain fit) -------------

def test_leg_kinematics(pkg, ob):
    cfg = pkg.prep_config_default()
    rng = np.random.default_rng(3)
    for leg in range(4):
        rho = np.array(cfg.rho_fix[5 * leg:5 * leg + 5])
        for _ in range(5):
            q = rng.uniform([-0.4, 0.3, -2.2], [0.4, 1.3, -1.0])
            p1, J1 = pkg.a1_leg_fk_jac(rho, q)          # product host function
            p2, J2 = ob.leg_fk_jac(rho, q)              # oracle, written from the rotation chain
            assert np.abs(p1 - p2).max() < 1e-15 and np.abs(J1 - J2).max() < 1e-15
            h = 1e-6
            Jfd = np.stack([(ob.leg_fk_jac(rho, q + h * np.eye(3)[k])[0] - ob.leg_fk_jac(rho, q - h * np.eye(3)[k])[0]) / (2 * h)
                            for k in range(3)], axis=1)
            assert np.abs(Jfd - J1).max() < 1e-9
            # chain identity: |foot - hip|^2 = d^2 + lt^2 + lc^2 + 2 lt lc cos(q2)
            hip = np.array([rho[0], rho[1], 0.0])
            assert abs(np.sum((p1 - hip) ** 2) - (rho[2] ** 2 + rho[3] ** 2 + rho[4] ** 2 + 2 * rho[3] * rho[4] * np.cos(q[2]))) < 1e-12
    # the standing pose of the reference's default foot position (config/gazebo_a1_mpc.yaml:17-31)
    p, _ = pkg.a1_leg_fk_jac(np.array(cfg.rho_fix[:5]), [0.0, 0.8, -1.6])
    assert abs(p[0] - 0.1881) < 0.02 and abs(p[1] - 0.12675) < 1e-6 and -0.32 < p[2] < -0.28


def test_prep_oracle_orientation_and_packing(pkg, ob):
    cfg = pkg.prep_config_default()
    N, T = 16, 3
    sens = np.stack([pkg.generate_sensors(1002, 0, N, t) for t in range(T)])
    st, tin, ex = ob.prep_stream(cfg, sens)
    ref = np.stack([pkg.generate_stream_states(1002, 0, N, t) for t in range(T)])
    # the sensor stream is the same robots seen through quaternion / IMU: the preparation recovers them
    assert np.abs(st["euler"] - ref["euler"]).max() < 1e-6
    assert np.abs(st["rot_mat"] - ref["rot_mat"]).max() < 1e-6
    assert np.abs(st["ang_vel"] - ref["ang_vel"]).max() < 1e-6
    for f in ("pos_d_z", "lin_vel_d", "ang_vel_d", "contacts"):
        assert np.array_equal(st[f], ref[f])
    # first tick: the estimator only initialises, root_pos stays the odometry value (GazeboA1ROS.cpp:194-198)
    assert np.array_equal(st["pos"][0], ref["pos"][0]) and np.array_equal(st["lin_vel"][0], ref["lin_vel"][0])
    assert not np.array_equal(st["pos"][1], ref["pos"][1])
    R = st["rot_mat"].reshape(T, N, 3, 3).astype(np.float64)
    prel = ex["foot_pos_rel"].reshape(T, N, 4, 3).astype(np.float64)
    assert np.abs(np.einsum("tnij,tnlj->tnli", R, prel).reshape(T, N, 12) - st["foot_pos_abs"]).max() < 1e-6
    J = tin["j_foot"].reshape(T, N, 4, 3, 3).astype(np.float64)
    qd = sens["joint_vel"].reshape(T, N, 4, 3).astype(np.float64)
    assert np.abs(np.einsum("tnlij,tnlj->tnli", J, qd).reshape(T, N, 12) - ex["foot_vel_rel"]).max() < 1e-5
    off = pkg.prep_config_default()
    off.use_estimator = 0
    off.use_terrain_adapt = 0
    st2, _, ex2 = ob.prep_stream(off, sens)
    assert np.array_equal(st2["pos"], ref["pos"]) and np.array_equal(st2["euler_d"], ref["euler_d"])


def test_prep_oracle_terrain_and_ekf(pkg, ob):
    cfg = pkg.prep_config_default()
    N, T = 3, 130
    base = pkg.generate_sensors(1002, 0, N, 0)
    # a standing robot on a known slope: constant sensors, feet exactly on z = -0.3 + 0.2 x - 0.05 y
    base["movement_mode"] = 0
    base["joint_vel"] = 0
    base["imu_ang_vel"] = 0
    base["root_quat"] = [1, 0, 0, 0]
    base["imu_acc"] = [0, 0, 9.81]
    base["root_pos"][:, 2] = 0.3
    frc = base["foot_pos_recent_contact"].reshape(N, 4, 3)
    frc[:, :, 2] = -0.3 + 0.2 * frc[:, :, 0] - 0.05 * frc[:, :, 1]
    frc[:, 0, 2] += 0.0  # FL+FR-RL-RR > 0.05 for slope 0.2 over 0.35 m
    sens = np.stack([base] * T)
    st, tin, ex = ob.prep_stream(cfg, sens)
    want = np.arccos(1.0 / np.sqrt(0.2 ** 2 + 0.05 ** 2 + 1.0))
    assert np.abs(ex["terrain_pitch_angle"][-1] - want).max() < 1e-6          # window of 100 is full
    assert np.abs(ex["terrain_pitch_angle"][49] - 0.5 * want).max() < 1e-6   # half full: the average divides by 100
    frd = frc[:, 0, 2] + frc[:, 1, 2] - frc[:, 2, 2] - frc[:, 3, 2]
    sign = np.where(frd > 0.05, -1.0, 1.0)
    assert np.abs(st["euler_d"][-1, :, 1] - sign * want).max() < 1e-6
    # estimator: a motionless robot whose feet are on the ground converges to height = -foot z
    h = -ex["foot_pos_rel"][-1].reshape(N, 4, 3)[:, :, 2].mean(axis=1)
    assert np.abs(ex["estimated_root_pos"][-1, :, 2] - h).max() < 5e-3
    assert np.abs(ex["estimated_root_vel"][-1]).max() < 5e-3


# ---- exact discretisation, drifting feet, gait-aware contacts (behind flags) ------

def test_exact_discretization_is_the_matrix_exponential(pkg, ob):
    from scipy.linalg import expm
    cfg = pkg.config_default()
    st = pkg.generate_states(1002, 0, 6)
    for i in range(6):
        A, B = ob.discretize_exact(cfg, st[i])
        inter = ob.mpc_build_intermediates(cfg, st[i])          # forward Euler: A_d = I + dt A_c, B_d = dt B_c
        Ac, Bc = (inter["A_d"] - np.eye(13)) / cfg.dt, inter["B_d"] / cfg.dt
        M = np.zeros((25, 25))
        M[:13, :13], M[:13, 13:] = Ac * cfg.dt, Bc * cfg.dt
        E = expm(M)
        assert np.abs(E[:13, :13] - A).max() < 1e-15 and np.abs(E[:13, 13:] - B).max() < 1e-15
        assert abs(A[5, 12] - 0.5 * cfg.dt ** 2) < 1e-18           # the one entry Euler misses in A_d
        assert np.abs(B[3:6] - 0.5 * cfg.dt * B[9:12]).max() < 1e-18


def test_extension_flags_off_is_the_reference_and_gait_schedule(pkg, ob):
    cfg = pkg.config_default()
    st = pkg.generate_states(1002, 0, 64)
    gait = pkg.generate_gait_inputs(1002, 0, 64, 0)
    P0, q0, l0, u0 = ob.mpc_build_qp(cfg, st[3])
    P1, q1, l1, u1 = ob.mpc_build_qp_ext(cfg, st[3], gait[3])    # flags 0: gait records are ignored
    assert np.array_equal(P0, P1) and np.array_equal(q0, q1) and np.array_equal(u0, u1)
    cfg.gait_aware = 1
    changed = 0
    for i in range(64):
        P, q, l, u = ob.mpc_build_qp_ext(cfg, st[i], gait[i])
        assert np.array_equal(P, ob.mpc_build_qp(pkg.config_default(), st[i])[0])   # only the bounds move
        # python restatement of the scheduler loop (A1RobotControl.cpp:156-164), one tick per MPC step
        cnt = gait["gait_counter"][i].astype(np.float64).copy()
        for h in range(10):
            if h > 0:
                cnt = np.fmod(cnt + gait["gait_counter_speed"][i], gait["counter_per_gait"][i])
            want = st["contacts"][i] != 0 if h == 0 else cnt <= gait["counter_per_swing"][i]
            assert np.array_equal(u[20 * h + 4:20 * h + 20:5] > 0, want), (i, h)
        changed += int(not np.array_equal(u, ob.mpc_build_qp(pkg.config_default(), st[i])[3]))
    assert changed >= 4   # some robots swap stance and swing inside the horizon
    cfg = pkg.config_default()
    cfg.foot_drift = 1
    moving = int(np.argmax(np.abs(st["lin_vel_d"]).sum(axis=1)))
    assert not np.array_equal(ob.mpc_build_qp_ext(cfg, st[moving])[0], ob.mpc_build_qp(pkg.config_default(), st[moving])[0])
