"""Prototype (numpy, CPU) of the wrench-space ADMM linear algebra that csrc/wrench_kernel.cuh runs.

The condensed MPC Hessian is  P = G' S G + R2  with
    G  = blockdiag_k(B6c_k)  (6H x 12H): inputs of step k -> net wrench w_k = [I_w^-1 sum r x f ; sum f / m]
    S  = calA' Q calA        (6H x 6H):  S_kl = sum_{i >= max(k,l)} Gam'(A^(i-k))' Q A^(i-l) Gam
    R2 = diag(2 r)
because B_d = Gam B6c (Gam = dt J for forward Euler; (dt I + A_c dt^2/2) J for the exact
discretisation), i.e. rank(B_qp) = 6H.  Then
    K = c D P D + sigma I + A_' rho A_ = G_' C G_ + Delta,   G_ = G D, C = c S,
    Delta = c D R2 D + sigma I + A_' rho A_   (block diagonal, 3x3 per leg-step)
and by Woodbury  K^-1 r = a - Delta^-1 G_' Y G_ a,  a = Delta^-1 r,
    Y = (C^-1 + N)^-1 = N^-1 - N^-1 (C + N^-1)^-1 N^-1,  N = G_ Delta^-1 G_'  (6x6 per step)
-- a 60 x 60 dense mat-vec per iteration instead of 120 x 120, a 60^3 factorisation instead of 120^3,
and no Hessian in memory at all.  This script checks that the iterates are the oracle's.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import osqp_independent as oi  # noqa: E402


def wrench_build(cfg, rec):
    """S (6H x 6H), gam (6H), B6c (H x 6 x 12) from one MpcStateIn record."""
    H = cfg.horizon
    dt = cfg.dt
    Qd = 2.0 * np.array(cfg.q_weights[:])
    euler = rec["euler"].astype(np.float64)
    Rm = rec["rot_mat"].astype(np.float64).reshape(3, 3)
    cy, sy = np.cos(euler[2]), np.sin(euler[2])
    Ac = np.zeros((13, 13))
    Ac[0:3, 6:9] = [[cy, sy, 0], [-sy, cy, 0], [0, 0, 1]]
    Ac[3:6, 9:12] = np.eye(3)
    Ac[11, 12] = 1
    Ad = np.eye(13) + dt * Ac
    I = np.array(cfg.inertia[:]).reshape(3, 3)
    Iw_inv = np.linalg.inv(Rm @ I @ Rm.T)
    foot = rec["foot_pos_abs"].astype(np.float64).reshape(4, 3)
    B6 = np.zeros((6, 12))
    for i in range(4):
        v = foot[i]
        sk = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        B6[0:3, 3 * i:3 * i + 3] = Iw_inv @ sk
        B6[3:6, 3 * i:3 * i + 3] = np.eye(3) / cfg.mass
    Gam = np.zeros((13, 6))
    Gam[6:12, :] = dt * np.eye(6)
    Apow = [np.eye(13)]
    for _ in range(H):
        Apow.append(Ad @ Apow[-1])
    AG = [Apow[m] @ Gam for m in range(H)]            # A^m Gam, 13 x 6
    S = np.zeros((6 * H, 6 * H))
    for k in range(H):
        for l in range(H):
            acc = np.zeros((6, 6))
            for i in range(max(k, l), H):
                acc += AG[i - k].T @ (Qd[:, None] * AG[i - l])
            S[6 * k:6 * k + 6, 6 * l:6 * l + 6] = acc
    x0 = np.concatenate([euler, rec["pos"], rec["ang_vel"], rec["lin_vel"], [-9.8]]).astype(np.float64)
    vdw = Rm @ rec["lin_vel_d"].astype(np.float64)
    gam = np.zeros(6 * H)
    e = []
    for i in range(H):
        xr = np.array([rec["euler_d"][0], rec["euler_d"][1], euler[2] + float(rec["ang_vel_d"][2]) * dt * (i + 1),
                       float(rec["pos"][0]) + vdw[0] * dt * (i + 1), float(rec["pos"][1]) + vdw[1] * dt * (i + 1),
                       rec["pos_d_z"], rec["ang_vel_d"][0], rec["ang_vel_d"][1], rec["ang_vel_d"][2],
                       vdw[0], vdw[1], 0.0, -9.8], dtype=np.float64)
        e.append(Apow[i + 1] @ x0 - xr)
    for k in range(H):
        acc = np.zeros(6)
        for i in range(k, H):
            acc += AG[i - k].T @ (Qd * e[i])
        gam[6 * k:6 * k + 6] = acc
    return S, gam, np.stack([B6] * H)


class WrenchOsqp(oi.Osqp):
    """The independent OSQP with its KKT solve replaced by the wrench-space Woodbury solve; the
    equilibration sees P only through G' S G + R2 evaluated entry by entry."""

    def __init__(self, S, gam, B6, r2, A, l, u, settings):
        self.S, self.B6, self.r2 = S, B6, r2
        H = B6.shape[0]
        self.H = H
        G = np.zeros((6 * H, 12 * H))
        for k in range(H):
            G[6 * k:6 * k + 6, 12 * k:12 * k + 12] = B6[k]
        self.G = G
        P = G.T @ S @ G + np.diag(r2)
        q = G.T @ gam
        super().__init__(P, q, A, l, u, settings)

    def _factor(self):
        n, H = self.n, self.H
        # Delta: block diagonal 3x3 per leg-step = diag(c D^2 r2 + sigma) + A_' rho A_ (A_ is the scaled A)
        Delta = np.diag(self.c * self.D ** 2 * self.r2 + self.s.sigma) + self.A.T @ (self.rho_v[:, None] * self.A)
        self.Dinv_blocks = np.zeros((n, n))
        for g in range(n // 3):
            sl = slice(3 * g, 3 * g + 3)
            self.Dinv_blocks[sl, sl] = np.linalg.inv(Delta[sl, sl])
        Gb = self.G * self.D[None, :]
        self.M1 = Gb @ self.Dinv_blocks                      # 6H x 12H, block diagonal per step
        N = self.M1 @ Gb.T                                   # block diagonal 6x6 per step
        Ninv = np.zeros_like(N)
        for k in range(H):
            sl = slice(6 * k, 6 * k + 6)
            Ninv[sl, sl] = np.linalg.inv(N[sl, sl])
        Z = np.linalg.inv(self.c * self.S + Ninv)
        self.Y = Ninv - Ninv @ Z @ Ninv
        self.Gb = Gb
        self.n_factor = getattr(self, "n_factor", 0) + 1
        self.lu = None

    def solve(self):
        # same Algorithm 1, linear system solved through Y
        import scipy.linalg as sla
        real = sla.lu_solve

        def wrench_solve(_lu, rhs, check_finite=False):
            n = self.n
            r = rhs[:n] + self.A.T @ (self.rho_v * rhs[n:])  # eliminate nu: K x~ = sigma x - q + A'(rho z - y)
            a = self.Dinv_blocks @ r
            xt = a - self.M1.T @ (self.Y @ (self.Gb @ a))
            zt = self.A @ xt
            nu = self.rho_v * (zt - rhs[n:])
            return np.concatenate([xt, nu])

        sla.lu_solve = wrench_solve
        try:
            return super().solve()
        finally:
            sla.lu_solve = real


def main():
    import go1_qp_mpc_controller_b200 as pkg
    import oracle_binding as ob
    from test_independent_osqp import _bounds
    from test_oracle import numpy_build
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    for name in ("gazebo", "hardware"):
        cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
        A = oi.mpc_constraint_matrix(10, cfg.mu)
        st = oi.Settings.from_ctypes(cfg.osqp)
        states = pkg.generate_states(1002, 0, n)
        ref = ob.mpc_compute_grf(cfg, states)
        r2 = np.tile(2.0 * np.array(cfg.r_weights[:]), 10)
        same = 0
        worst = 0.0
        worstP = 0.0
        for i in range(n):
            S, gam, B6 = wrench_build(cfg, states[i])
            l, u = _bounds(cfg, states[i])
            s = WrenchOsqp(S, gam, B6, r2, A, l, u, st)
            if i < 4:
                b = numpy_build(cfg, states[i])
                Pw = s.G.T @ S @ s.G + np.diag(r2)
                worstP = max(worstP, np.abs(Pw - b["P"]).max() / np.abs(b["P"]).max(),
                             np.abs(s.G.T @ gam - b["q"]).max() / np.abs(b["q"]).max())
            s.solve()
            g = oi.grf_body(s.solution(), states[i]["rot_mat"])
            ok = s.iters == ref["iters"][i] and s.rho_updates == ref["rho_updates"][i]
            same += ok
            if ok:
                worst = max(worst, np.linalg.norm(g - ref["grf"][i]) / max(np.linalg.norm(ref["grf"][i]), 1.0))
        print(f"{name}: same iterate sequence {same}/{n}, max GRF rel err {worst:.2e}, P/q rel err {worstP:.2e}")


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------
# Second stage of the prototype: the formulas exactly as csrc/wrench_kernel.cuh evaluates them.
#   * closed-form S and gamma (A_c is nilpotent: F_m = A_d^m Gam has two non-zeros per column)
#   * the factorisation through the block Cholesky factor of N ("L-route"): no inverse of N or C,
#     the swept matrix I + L' C L has eigenvalues >= 1
# ---------------------------------------------------------------------------------------------
def wrench_build_closed_form(cfg, rec, exact=False):
    H, dt = cfg.horizon, cfg.dt
    Q = 2.0 * np.array(cfg.q_weights[:])
    euler = rec["euler"].astype(np.float64)
    Rm = rec["rot_mat"].astype(np.float64).reshape(3, 3)
    cy, sy = np.cos(euler[2]), np.sin(euler[2])
    RzT = np.array([[cy, sy, 0], [-sy, cy, 0], [0, 0, 1.0]])        # A_c block (0, 6)
    Theta = RzT.T @ np.diag(Q[0:3]) @ RzT                             # sum_r RzT[r,c] Q_r RzT[r,c~]
    D1 = np.diag(np.concatenate([Q[6:9], Q[9:12]]))
    D2 = np.zeros((6, 6))
    D2[0:3, 0:3] = Theta
    D2[3:6, 3:6] = np.diag(Q[3:6])
    kap = np.arange(H) + (0.5 if exact else 0.0)
    S = np.zeros((6 * H, 6 * H))
    for k in range(H):
        for l in range(H):
            mx = max(k, l)
            alpha = (H - mx) * dt * dt
            beta = dt ** 4 * sum(kap[i - k] * kap[i - l] for i in range(mx, H))
            S[6 * k:6 * k + 6, 6 * l:6 * l + 6] = alpha * D1 + beta * D2
    x0 = np.concatenate([euler, rec["pos"], rec["ang_vel"], rec["lin_vel"], [-9.8]]).astype(np.float64)
    vdw = Rm @ rec["lin_vel_d"].astype(np.float64)
    Acx = np.zeros(13)
    Acx[0:3] = RzT @ x0[6:9]
    Acx[3:6] = x0[9:12]
    Acx[11] = x0[12]
    Qe = np.zeros((H, 13))
    for i in range(H):
        m = i + 1
        c2 = (m * dt) ** 2 / 2 if exact else m * (m - 1) / 2 * dt * dt
        xi = x0 + m * dt * Acx
        xi[5] += c2 * x0[12]
        xr = np.array([rec["euler_d"][0], rec["euler_d"][1], euler[2] + float(rec["ang_vel_d"][2]) * dt * m,
                       float(rec["pos"][0]) + vdw[0] * dt * m, float(rec["pos"][1]) + vdw[1] * dt * m,
                       rec["pos_d_z"], rec["ang_vel_d"][0], rec["ang_vel_d"][1], rec["ang_vel_d"][2],
                       vdw[0], vdw[1], 0.0, -9.8], dtype=np.float64)
        Qe[i] = Q * (xi - xr)
    gam = np.zeros(6 * H)
    for k in range(H):
        for i in range(k, H):
            kp = kap[i - k]
            gam[6 * k:6 * k + 3] += dt * Qe[i, 6:9] + kp * dt * dt * (RzT.T @ Qe[i, 0:3])
            gam[6 * k + 3:6 * k + 6] += dt * Qe[i, 9:12] + kp * dt * dt * Qe[i, 3:6]
    return S, gam


class WrenchOsqpL(WrenchOsqp):
    def _factor(self):
        n, H = self.n, self.H
        Delta = np.diag(self.c * self.D ** 2 * self.r2 + self.s.sigma) + self.A.T @ (self.rho_v[:, None] * self.A)
        Di = np.zeros((n, n))
        for g in range(n // 3):
            sl = slice(3 * g, 3 * g + 3)
            Di[sl, sl] = np.linalg.inv(Delta[sl, sl])
        Gb = self.G * self.D[None, :]
        M1 = Gb @ Di
        N = M1 @ Gb.T
        L = np.zeros_like(N)
        for k in range(H):
            sl = slice(6 * k, 6 * k + 6)
            L[sl, sl] = np.linalg.cholesky(N[sl, sl])
        Linv = np.linalg.inv(L)
        Gh = Linv @ Gb                      # G^ : rows orthonormal in the Delta^-1 metric
        Mh = Linv @ M1                      # = G^ Delta^-1
        W = L.T @ (self.c * self.S) @ L
        Yp = np.eye(6 * H) - np.linalg.inv(np.eye(6 * H) + W)
        # same operator as the N^-1 route
        self.Dinv_blocks, self.Gb, self.M1, self.Y = Di, Gh, Mh, Yp
        self.n_factor = getattr(self, "n_factor", 0) + 1
        self.lu = None


def main2():
    import go1_qp_mpc_controller_b200 as pkg
    import oracle_binding as ob
    from test_independent_osqp import _bounds
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    for name in ("gazebo", "hardware"):
        cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
        A = oi.mpc_constraint_matrix(10, cfg.mu)
        st = oi.Settings.from_ctypes(cfg.osqp)
        states = pkg.generate_states(1003, 0, n)
        ref = ob.mpc_compute_grf(cfg, states)
        r2 = np.tile(2.0 * np.array(cfg.r_weights[:]), 10)
        same, worst, worstS = 0, 0.0, 0.0
        for i in range(n):
            S0, gam0, B6 = wrench_build(cfg, states[i])
            S, gam = wrench_build_closed_form(cfg, states[i])
            worstS = max(worstS, np.abs(S - S0).max() / np.abs(S0).max(), np.abs(gam - gam0).max() / np.abs(gam0).max())
            l, u = _bounds(cfg, states[i])
            s = WrenchOsqpL(S, gam, B6, r2, A, l, u, st).solve()
            g = oi.grf_body(s.solution(), states[i]["rot_mat"])
            ok = s.iters == ref["iters"][i] and s.rho_updates == ref["rho_updates"][i]
            same += ok
            if ok:
                worst = max(worst, np.linalg.norm(g - ref["grf"][i]) / max(np.linalg.norm(ref["grf"][i]), 1.0))
        print(f"[closed form + L-route] {name}: same iterate sequence {same}/{n}, max GRF rel err {worst:.2e}, "
              f"S/gamma closed-form rel err {worstS:.2e}")


if __name__ == "__main__" and len(sys.argv) > 2:
    main2()
