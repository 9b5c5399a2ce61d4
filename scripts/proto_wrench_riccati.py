"""Prototype (numpy, CPU) of the long-horizon linear algebra of csrc/wrench_riccati_kernel.cuh.

scripts/proto_wrench.py established  K = G_' C G_ + Delta  (G_ = G D block diagonal 6 x 12 per step,
C = c S the 6H x 6H wrench-space Hessian, Delta block diagonal 3 x 3 per leg-step) and
    K^-1 r = a - M~' (tau - dlt),   a = Delta^-1 r,  tau = G_ a,  M~ = N^-1 G_ Delta^-1,  N = G_ Delta^-1 G_',
    dlt = (C + N^-1)^-1 N^-1 tau.
At H = 10 the 60 x 60 core is dense (registers).  At H = 30 it is the Hessian of an LQR problem with SIX
inputs per step -- the velocity increment dlt_k = dt B6c_k D_k u_k of the step -- and is solved exactly by a
Riccati recursion on the 12 states (the gravity state never moves in a Hessian solve):
    state X = (pos 6 = euler, position; vel 6 = angular, linear),  A = [[I, dt Rt], [0, I]],  Rt = blockdiag(RzT, I)
    X_k+1 = A X_k + Gam dlt_k,  Gam = [Gp; I]  (Gp = dt/2 Rt with exact_discretization, else 0)
    cost  sum_k 1/2 X_k+1' cQ X_k+1 + 1/2 dlt_k' N_k^-1 dlt_k - u_k' dlt_k,   u = N^-1 tau = M~ r
  factor:  Pi = cQ;  k = H-1 .. 0:  T = Pi Gam, Z_k = (N_k^-1 + Gam' T)^-1 = L (I + L' Gam' T L)^-1 L'  (N = L L'),
           U = T' A,  F_k = -Z_k U,  Pi <- cQ + A' Pi A + U' F_k
  solve:   p_H = 0;  e_k = Gam' p_k+1 - u_k,  p_k = A' p_k+1 + F_k' e_k          (backward)
           X_0 = 0;  dlt_k = F_k X_k - Z_k e_k,  X_k+1 = A X_k + Gam dlt_k          (forward)
No inverse of N is formed except through its Cholesky factor; the pivots of I + L' (.) L are >= 1.
This script checks the iterates against the oracle at H = 30 (identical iteration counts).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import osqp_independent as oi  # noqa: E402
import proto_wrench as pw  # noqa: E402


class WrenchRiccati(pw.WrenchOsqp):
    """WrenchOsqp with the core solve replaced by the 12-state / 6-input Riccati recursion."""

    def __init__(self, S, gam, B6, r2, A, l, u, settings, dt, RzT, Qd, exact=False, group=4):
        self.dt, self.RzT, self.Qd, self.exact, self.group = dt, RzT, Qd, exact, group
        super().__init__(S, gam, B6, r2, A, l, u, settings)

    def _factor(self):
        n, H, dt = self.n, self.H, self.dt
        Delta = np.diag(self.c * self.D ** 2 * self.r2 + self.s.sigma) + self.A.T @ (self.rho_v[:, None] * self.A)
        Di = np.zeros((n, n))
        for g in range(n // 3):
            sl = slice(3 * g, 3 * g + 3)
            Di[sl, sl] = np.linalg.inv(Delta[sl, sl])
        self.Di = Di
        Rt = np.zeros((6, 6))
        Rt[0:3, 0:3] = self.RzT
        Rt[3:6, 3:6] = np.eye(3)
        Amat = np.eye(12)
        Amat[0:6, 6:12] = dt * Rt
        Gam = np.zeros((12, 6))
        Gam[6:12] = np.eye(6)
        if self.exact:
            Gam[0:6] = 0.5 * dt * Rt
        cQ = self.c * np.diag(np.concatenate([self.Qd[0:6], self.Qd[6:12]]))
        self.Amat, self.Gam = Amat, Gam
        self.Gd, self.Mt, self.Nk, self.Lk = [], [], [], []
        for k in range(H):
            Gd = dt * self.B6[k] * self.D[None, 12 * k:12 * k + 12]          # dlt = Gd u  (velocity increment)
            M1 = Gd @ Di[12 * k:12 * k + 12, 12 * k:12 * k + 12]
            N = M1 @ Gd.T
            L = np.linalg.cholesky(N)
            Mt = np.linalg.solve(L.T, np.linalg.solve(L, M1))               # N^-1 M1 through the factor
            self.Gd.append(Gd); self.Mt.append(Mt); self.Nk.append(N); self.Lk.append(L)
        Pi = cQ.copy()
        self.F, self.Z = [None] * H, [None] * H
        for k in range(H - 1, -1, -1):
            L = self.Lk[k]
            T = Pi @ Gam
            W = L.T @ (Gam.T @ T) @ L
            Z = L @ np.linalg.inv(np.eye(6) + W) @ L.T
            U = T.T @ Amat
            F = -Z @ U
            Pi = cQ + Amat.T @ Pi @ Amat + U.T @ F
            Pi = 0.5 * (Pi + Pi.T)
            self.F[k], self.Z[k] = F, Z
        # group transitions of the three-sweep recursion (the kernel's critical path): Phi_j = Acl_(gj+g-1) ... Acl_gj
        g = self.group
        self.Acl = [Amat + Gam @ self.F[k] for k in range(H)]
        self.Phi = []
        for j in range((H + g - 1) // g):
            M = np.eye(12)
            for k in range(g * j, min(g * j + g, H)):
                M = self.Acl[k] @ M
            self.Phi.append(M)
        self.n_factor = getattr(self, "n_factor", 0) + 1
        self.lu = None

    def core_solve(self, r):
        """x~ = K^-1 r."""
        H = self.H
        a = self.Di @ r
        u = [self.Mt[k] @ r[12 * k:12 * k + 12] for k in range(H)]
        tau = [self.Nk[k] @ u[k] for k in range(H)]                          # = G_ a
        A, Gam, g = self.Amat, self.Gam, self.group
        nG = (H + g - 1) // g
        # backward: three sweeps (groups from a zero boundary, boundaries through Phi', groups again)
        p = np.zeros((H + 1, 12))
        e = np.zeros((H, 6))

        def bwd_group(j, boundary):
            hi = min(g * j + g, H)
            p[hi] = boundary
            for k in range(hi - 1, g * j - 1, -1):
                e[k] = Gam.T @ p[k + 1] - u[k]
                p[k] = A.T @ p[k + 1] + self.F[k].T @ e[k]

        for j in range(nG):
            bwd_group(j, np.zeros(12))
        p0 = [p[g * j].copy() for j in range(nG)]
        pb = [np.zeros(12) for _ in range(nG + 1)]
        for j in range(nG - 1, -1, -1):
            pb[j] = self.Phi[j].T @ pb[j + 1] + p0[j]
        for j in range(nG):
            bwd_group(j, pb[j + 1])
        # forward
        X = np.zeros((H + 1, 12))
        dl = np.zeros((H, 6))
        b = [-self.Z[k] @ e[k] for k in range(H)]

        def fwd_group(j, boundary):
            X[g * j] = boundary
            for k in range(g * j, min(g * j + g, H)):
                dl[k] = self.F[k] @ X[k] + b[k]
                X[k + 1] = A @ X[k] + Gam @ dl[k]

        x0 = []
        for j in range(nG):
            fwd_group(j, np.zeros(12))
            x0.append(X[min(g * j + g, H)].copy())
        xb = [np.zeros(12)]
        for j in range(nG):
            xb.append(self.Phi[j] @ xb[j] + x0[j])
        for j in range(nG):
            fwd_group(j, xb[j])
        xt = a.copy()
        for k in range(H):
            xt[12 * k:12 * k + 12] -= self.Mt[k].T @ (tau[k] - dl[k])
        return xt

    def solve(self):
        import scipy.linalg as sla
        real = sla.lu_solve

        def wrench_solve(_lu, rhs, check_finite=False):
            n = self.n
            r = rhs[:n] + self.A.T @ (self.rho_v * rhs[n:])
            xt = self.core_solve(r)
            zt = self.A @ xt
            nu = self.rho_v * (zt - rhs[n:])
            return np.concatenate([xt, nu])

        sla.lu_solve = wrench_solve
        try:
            return oi.Osqp.solve(self)
        finally:
            sla.lu_solve = real


def main():
    import go1_qp_mpc_controller_b200 as pkg
    import oracle_binding as ob
    from test_independent_osqp import _bounds
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    for name in ("gazebo", "hardware"):
        cfg = pkg.config_default() if name == "gazebo" else pkg.config_hardware()
        cfg.horizon = H
        A = oi.mpc_constraint_matrix(H, cfg.mu)
        st = oi.Settings.from_ctypes(cfg.osqp)
        states = pkg.generate_states(1004, 0, n)
        ref = ob.mpc_compute_grf(cfg, states)
        r2 = np.tile(2.0 * np.array(cfg.r_weights[:]), H)
        Qd = 2.0 * np.array(cfg.q_weights[:])
        same, worst, worstK = 0, 0.0, 0.0
        for i in range(n):
            S, gam = pw.wrench_build_closed_form(cfg, states[i])
            _, _, B6 = pw.wrench_build(cfg, states[i]) if H <= 10 else (None, None, None)
            if B6 is None:
                B6 = b6_of(cfg, states[i], H)
            l, u = _bounds(cfg, states[i])
            yaw = float(states[i]["euler"][2])
            cy, sy = np.cos(yaw), np.sin(yaw)
            RzT = np.array([[cy, sy, 0], [-sy, cy, 0], [0, 0, 1.0]])
            s = WrenchRiccati(S, gam, B6, r2, A, l, u, st, cfg.dt, RzT, Qd)
            # the structured solve against the dense reduced matrix, on a random right-hand side
            rng = np.random.default_rng(i)
            r = rng.standard_normal(s.n)
            K = s.P + st.sigma * np.eye(s.n) + s.A.T @ (s.rho_v[:, None] * s.A)
            xd = np.linalg.solve(K, r)
            worstK = max(worstK, np.abs(s.core_solve(r) - xd).max() / np.abs(xd).max())
            s.solve()
            g = oi.grf_body(s.solution(), states[i]["rot_mat"])
            ok = s.iters == ref["iters"][i] and s.rho_updates == ref["rho_updates"][i]
            same += ok
            if ok:
                worst = max(worst, np.linalg.norm(g - ref["grf"][i]) / max(np.linalg.norm(ref["grf"][i]), 1.0))
            print(f"  state {i}: iters {s.iters} (oracle {ref['iters'][i]}), rho updates {s.rho_updates} "
                  f"({ref['rho_updates'][i]}), factorisations {s.n_factor}", flush=True)
        print(f"{name} H={H}: same iterate sequence {same}/{n}, max GRF rel err {worst:.2e}, "
              f"core solve vs dense rel err {worstK:.2e}")


def b6_of(cfg, rec, H):
    Rm = rec["rot_mat"].astype(np.float64).reshape(3, 3)
    I = np.array(cfg.inertia[:]).reshape(3, 3)
    Iw_inv = np.linalg.inv(Rm @ I @ Rm.T)
    foot = rec["foot_pos_abs"].astype(np.float64).reshape(4, 3)
    B6 = np.zeros((6, 12))
    for i in range(4):
        v = foot[i]
        sk = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        B6[0:3, 3 * i:3 * i + 3] = Iw_inv @ sk
        B6[3:6, 3 * i:3 * i + 3] = np.eye(3) / cfg.mass
    return np.stack([B6] * H)


if __name__ == "__main__":
    main()
