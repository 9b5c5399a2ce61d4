"""An INDEPENDENT second implementation of the OSQP 0.6.x iteration -- TEST INFRASTRUCTURE.

Purpose: pin `oracle/` (the C++ restatement the GPU is compared with).  The reference's solve
(A1RobotControl.cpp:416-439, :522-555; test/test_mpc.cpp:131-151) runs in un-vendored OSQP 0.6.x
through OsqpEigen 0.6.3; neither is in this image and the reference stores no expected output.
This file is written from the algorithm as published (Stellato, Banjac, Goulart, Bemporad, Boyd:
"OSQP: an operator splitting solver for quadratic programs", Math. Prog. Comp. 12, 2020;
Algorithm 1, section 3.1 "solving the linear system", 5.1 "preconditioning", 5.2 "parameter
selection", 3.4 termination) and the library's documented settings, NOT from `oracle/`:

  * the linear system of every iteration is the full quasi-definite KKT system
        [ P + sigma I    A'          ] [ x~ ]   [ sigma x_k - q        ]
        [ A             -diag(rho)^-1 ] [ nu ] = [ z_k - diag(rho)^-1 y ]     z~ = z_k + rho^-1 (nu - y)
    factorised densely with LAPACK LU (partial pivoting).  The oracle and the GPU both work with
    the REDUCED matrix K = P + sigma I + A' diag(rho) A (Cholesky / explicit inverse); the two
    are equal in exact arithmetic only, so agreement of iteration counts is evidence, not a tautology;
  * numpy vector code, dense A, no shared helper, different operation order everywhere.

What it deliberately shares with the oracle: the SETTINGS (the reference changes none but verbose
and warm_start, A1RobotControl.cpp:523-524; the adaptive-rho interval is pinned because the
library default depends on wall-clock time).
"""
import numpy as np
import scipy.linalg as sla

OSQP_INFTY = 1e30          # OsqpEigen::INFTY (ConvexMpc.cpp:229-237)
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_TOL = 1e-4             # l and u closer than this: equality row
RHO_EQ_OVER_RHO_INEQ = 1e3

SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED = 1, 2, -2
PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, UNSOLVED = -3, -4, -10


class Settings:
    def __init__(self, **kw):
        self.rho = 0.1
        self.sigma = 1e-6
        self.alpha = 1.6
        self.eps_abs = 1e-3
        self.eps_rel = 1e-3
        self.eps_prim_inf = 1e-4
        self.eps_dual_inf = 1e-4
        self.max_iter = 4000
        self.check_termination = 25
        self.scaling = 10
        self.adaptive_rho = 1
        self.adaptive_rho_interval = 50   # pinned (library default 0 = wall-clock dependent)
        self.adaptive_rho_tolerance = 5.0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    @classmethod
    def from_ctypes(cls, s):
        return cls(**{k: getattr(s, k) for k in ("rho", "sigma", "alpha", "eps_abs", "eps_rel", "eps_prim_inf",
                                                  "eps_dual_inf", "max_iter", "check_termination", "scaling",
                                                  "adaptive_rho", "adaptive_rho_interval", "adaptive_rho_tolerance")})


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.minimum(v, MAX_SCALING)


def mpc_constraint_matrix(H, mu):
    """linear_constraints of ConvexMpc.cpp:46-58: per leg-step rows fx+mu fz, fx-mu fz, fy+mu fz, fy-mu fz, fz."""
    A = np.zeros((20 * H, 12 * H))
    for k in range(4 * H):
        r, c = 5 * k, 3 * k
        A[r + 0, c + 0] = A[r + 1, c + 0] = 1.0
        A[r + 2, c + 1] = A[r + 3, c + 1] = 1.0
        A[r + 4, c + 2] = 1.0
        A[r + 0, c + 2] = A[r + 2, c + 2] = mu
        A[r + 1, c + 2] = A[r + 3, c + 2] = -mu
    return A


def balance_constraint_matrix(mu):
    """A1RobotControl.cpp:28-48: rows 0-3 select fz_i; then per leg +-fx - mu fz, +-fy - mu fz."""
    A = np.zeros((20, 12))
    for i in range(4):
        A[i, 2 + 3 * i] = 1.0
        A[4 + 4 * i, 3 * i] = 1.0
        A[5 + 4 * i, 3 * i] = -1.0
        A[6 + 4 * i, 1 + 3 * i] = 1.0
        A[7 + 4 * i, 1 + 3 * i] = -1.0
        A[4 + 4 * i:8 + 4 * i, 2 + 3 * i] = -mu
    return A


class Osqp:
    """osqp_setup / osqp_solve / osqp_update_{P, lin_cost, bounds} on dense data."""

    def __init__(self, P, q, A, l, u, settings):
        self.s = settings
        self.n, self.m = len(q), len(l)
        self.A_orig = np.array(A, dtype=np.float64)
        self.rho = float(settings.rho)
        self._load(P, q, l, u)
        self._scale()
        self.ctype = np.full(self.m, 99)
        self._rho_vec()
        self._factor()
        self.x = np.zeros(self.n)
        self.z = np.zeros(self.m)
        self.y = np.zeros(self.m)
        self._reset_info()

    # ---- data -----------------------------------------------------------------------------
    def _load(self, P, q, l, u):
        P = np.array(P, dtype=np.float64)
        # OsqpEigen hands OSQP the upper triangle only; the solver works with its symmetric completion
        self.P = np.triu(P) + np.triu(P, 1).T
        self.q = np.array(q, dtype=np.float64)
        self.A = self.A_orig.copy()
        self.l = np.array(l, dtype=np.float64)
        self.u = np.array(u, dtype=np.float64)

    def _scale(self):
        """Section 5.1, Algorithm 2 (modified Ruiz equilibration) as the library runs it."""
        n, m = self.n, self.m
        self.D, self.E, self.c = np.ones(n), np.ones(m), 1.0
        for _ in range(self.s.scaling):
            col_kkt_top = np.maximum(np.abs(self.P).max(axis=0), np.abs(self.A).max(axis=0)) if m else np.abs(self.P).max(axis=0)
            col_kkt_bot = np.abs(self.A).max(axis=1) if m else np.zeros(0)
            d = 1.0 / np.sqrt(_limit(col_kkt_top))
            e = 1.0 / np.sqrt(_limit(col_kkt_bot))
            self.P = d[:, None] * self.P * d[None, :]
            self.A = e[:, None] * self.A * d[None, :]
            self.q = d * self.q
            self.D *= d
            self.E *= e
            # cost normalisation
            mean_col = np.abs(self.P).max(axis=0).mean()
            qn = float(_limit(np.array(np.abs(self.q).max())))
            g = 1.0 / float(_limit(np.array(max(mean_col, qn))))
            self.P *= g
            self.q *= g
            self.c *= g
        self.l = self.E * self.l
        self.u = self.E * self.u

    def _rho_vec(self):
        lo_inf = self.l < -OSQP_INFTY * MIN_SCALING
        up_inf = self.u > OSQP_INFTY * MIN_SCALING
        ctype = np.zeros(self.m, dtype=int)
        ctype[lo_inf & up_inf] = -1
        ctype[~(lo_inf & up_inf) & (self.u - self.l < RHO_TOL)] = 1
        changed = not np.array_equal(ctype, self.ctype)
        self.ctype = ctype
        self.rho_v = np.where(ctype == -1, RHO_MIN, np.where(ctype == 1, RHO_EQ_OVER_RHO_INEQ * self.rho, self.rho))
        return changed

    def _factor(self):
        n, m = self.n, self.m
        K = np.empty((n + m, n + m))
        K[:n, :n] = self.P + self.s.sigma * np.eye(n)
        K[:n, n:] = self.A.T
        K[n:, :n] = self.A
        K[n:, n:] = -np.diag(1.0 / self.rho_v)
        self.lu = sla.lu_factor(K, check_finite=False)
        self.n_factor = getattr(self, "n_factor", 0) + 1

    def _reset_info(self):
        self.status = UNSOLVED
        self.iters = 0
        self.rho_updates = 0
        self.pri_res = self.dua_res = 0.0

    # ---- updates of a live solver (the reference's warm path, A1RobotControl.cpp:532-538) ----
    def update(self, P, q, l, u):
        """OsqpEigen updateHessianMatrix, updateGradient, updateLowerBound, updateUpperBound, in the
        reference's order.  osqp_update_P un-scales the data, installs P, re-runs the equilibration
        from scratch (with the gradient and bounds of the PREVIOUS solve still in place) and
        refactors; the other three replace and scale their vector; bound updates re-type rho_vec."""
        q_old = self.q / (self.c * self.D)
        l_old, u_old = self.l / self.E, self.u / self.E
        self._load(P, q_old, l_old, u_old)
        self._scale()
        self._factor()
        self._reset_info()
        self.q = self.c * self.D * np.array(q, dtype=np.float64)
        self.l = self.E * np.array(l, dtype=np.float64)
        if (self.l > self.u).any():
            raise ValueError("lower bound above the previous upper bound")
        if self._rho_vec():
            self._factor()
        self.u = self.E * np.array(u, dtype=np.float64)
        if (self.l > self.u).any():
            raise ValueError("upper bound below lower bound")
        if self._rho_vec():
            self._factor()

    # ---- residuals and certificates (section 3.4) ---------------------------------------------
    def _residuals(self):
        Ax = self.A @ self.x
        Px = self.P @ self.x
        Aty = self.A.T @ self.y
        rp = Ax - self.z
        rd = Px + self.q + Aty
        Ei, Di, ci = 1.0 / self.E, 1.0 / self.D, 1.0 / self.c
        r = {
            "pri": np.abs(Ei * rp).max() if self.m else 0.0,
            "dua": ci * np.abs(Di * rd).max(),
            "pri_sc": np.abs(rp).max() if self.m else 0.0,
            "dua_sc": np.abs(rd).max(),
            "pri_nrm": max(np.abs(Ei * self.z).max(), np.abs(Ei * Ax).max()) if self.m else 0.0,
            "dua_nrm": ci * max(np.abs(Di * self.q).max(), np.abs(Di * Aty).max(), np.abs(Di * Px).max()),
            "pri_nrm_sc": max(np.abs(self.z).max(), np.abs(Ax).max()) if self.m else 0.0,
            "dua_nrm_sc": max(np.abs(self.q).max(), np.abs(Aty).max(), np.abs(Px).max()),
        }
        self.pri_res, self.dua_res = float(r["pri"]), float(r["dua"])
        return r

    def _primal_infeasible(self, eps):
        dy = self.dy.copy()
        lo_inf = self.l < -OSQP_INFTY * MIN_SCALING
        up_inf = self.u > OSQP_INFTY * MIN_SCALING
        dy[lo_inf & up_inf] = 0.0
        dy[up_inf & ~lo_inf] = np.minimum(dy[up_inf & ~lo_inf], 0.0)
        dy[lo_inf & ~up_inf] = np.maximum(dy[lo_inf & ~up_inf], 0.0)
        nrm = np.abs(self.E * dy).max()
        if not nrm > eps:
            return False
        support = (self.u * np.maximum(dy, 0.0) + self.l * np.minimum(dy, 0.0)).sum()
        if not support < -eps * nrm:
            return False
        return np.abs((self.A.T @ dy) / self.D).max() < eps * nrm

    def _dual_infeasible(self, eps):
        dx = self.dx
        nrm = np.abs(self.D * dx).max()
        if not nrm > eps:
            return False
        if not (self.q @ dx) / self.c < -eps * nrm:
            return False
        if not np.abs((self.P @ dx) / self.D).max() / self.c < eps * nrm:
            return False
        Adx = (self.A @ dx) / self.E
        lo_inf = self.l < -OSQP_INFTY * MIN_SCALING
        up_inf = self.u > OSQP_INFTY * MIN_SCALING
        bad = (~up_inf & (Adx > eps * nrm)) | (~lo_inf & (Adx < -eps * nrm))
        return not bad.any()

    def _terminated(self, r, approximate):
        k = 10.0 if approximate else 1.0
        eps_p = k * (self.s.eps_abs + self.s.eps_rel * r["pri_nrm"])
        eps_d = k * (self.s.eps_abs + self.s.eps_rel * r["dua_nrm"])
        p_ok = self.m == 0 or r["pri"] < eps_p
        d_ok = r["dua"] < eps_d
        p_inf = (not p_ok) and self._primal_infeasible(k * self.s.eps_prim_inf)
        d_inf = (not d_ok) and self._dual_infeasible(k * self.s.eps_dual_inf)
        if p_ok and d_ok:
            self.status = SOLVED_INACCURATE if approximate else SOLVED
        elif p_inf:
            self.status = 3 if approximate else PRIMAL_INFEASIBLE
        elif d_inf:
            self.status = 4 if approximate else DUAL_INFEASIBLE
        else:
            return False
        return True

    # ---- Algorithm 1 ----------------------------------------------------------------------
    def solve(self):
        s = self.s
        n = self.n
        self.dx, self.dy = np.zeros(n), np.zeros(self.m)
        it, checked = 0, False
        for it in range(1, s.max_iter + 1):
            rho_inv = 1.0 / self.rho_v
            rhs = np.concatenate([s.sigma * self.x - self.q, self.z - rho_inv * self.y])
            sol = sla.lu_solve(self.lu, rhs, check_finite=False)
            xt = sol[:n]
            zt = self.z + rho_inv * (sol[n:] - self.y)
            x_new = s.alpha * xt + (1.0 - s.alpha) * self.x
            self.dx = x_new - self.x
            self.x = x_new
            zr = s.alpha * zt + (1.0 - s.alpha) * self.z
            self.z = np.clip(zr + rho_inv * self.y, self.l, self.u)
            self.dy = self.rho_v * (zr - self.z)
            self.y = self.y + self.dy
            checked = False
            can_check = s.check_termination and it % s.check_termination == 0
            can_adapt = s.adaptive_rho and s.adaptive_rho_interval and it % s.adaptive_rho_interval == 0
            if can_check or can_adapt:
                r = self._residuals()
            if can_check:
                checked = True
                if self._terminated(r, False):
                    break
            if can_adapt:
                pn = r["pri_sc"] / (r["pri_nrm_sc"] + 1e-10)
                dn = r["dua_sc"] / (r["dua_nrm_sc"] + 1e-10)
                est = min(max(self.rho * np.sqrt(pn / (dn + 1e-10)), RHO_MIN), RHO_MAX)
                if est > self.rho * s.adaptive_rho_tolerance or est < self.rho / s.adaptive_rho_tolerance:
                    self.rho = float(est)
                    self.rho_v = np.where(self.ctype == -1, RHO_MIN,
                                          np.where(self.ctype == 1, RHO_EQ_OVER_RHO_INEQ * self.rho, self.rho))
                    self._factor()
                    self.rho_updates += 1
        self.iters = it
        if not checked:
            self._terminated(self._residuals(), False)
        if self.status == UNSOLVED:
            if not self._terminated(self._residuals(), True):
                self.status = MAX_ITER_REACHED
        return self

    def solution(self):
        ok = self.status in (SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED)
        return self.D * self.x if ok else np.full(self.n, np.nan)


def grf_body(x, rot_mat, legs=4):
    """First-step forces rotated to the body frame, R' f, zeros on NaN (A1RobotControl.cpp:555-561)."""
    R = np.asarray(rot_mat, dtype=np.float64).reshape(3, 3)
    out = np.zeros(3 * legs)
    for i in range(legs):
        f = x[3 * i:3 * i + 3]
        out[3 * i:3 * i + 3] = 0.0 if np.isnan(f).any() else R.T @ f
    return out
