"""Can the iteration count of a cold H = 10 solve be predicted while it runs?  (CPU only: the independent OSQP
restatement of tests/, traces of the termination ratios at every check.)  Context: the 4096-state launch loses
~8 % to an uneven end of the persistent CTAs; a longest-first order needs a predictor."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402
import oracle_binding as ob  # noqa: E402
import osqp_independent as oi  # noqa: E402


class Traced(oi.Osqp):
    def _terminated(self, r, approximate):
        eps_p = self.s.eps_abs + self.s.eps_rel * r["pri_nrm"]
        eps_d = self.s.eps_abs + self.s.eps_rel * r["dua_nrm"]
        self.trace.append((r["pri"] / eps_p, r["dua"] / eps_d, self.rho))
        return super()._terminated(r, approximate)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    cfg = pkg.config_default()
    H = cfg.horizon
    A = oi.mpc_constraint_matrix(H, cfg.mu)
    st = oi.Settings.from_ctypes(cfg.osqp)
    states = pkg.generate_states(1234, 0, n)
    rows = []
    for i in range(n):
        P, q, l, u = ob.mpc_build_qp(cfg, states[i])
        s = Traced(P, q, A, l, u, st)
        s.trace = []
        s.solve()
        rows.append((s.iters, s.trace))
    its = np.array([r[0] for r in rows], float)
    print(f"{n} states: iterations mean {its.mean():.1f} sd {its.std():.1f} min {its.min():.0f} max {its.max():.0f}")
    for at in (2, 4, 6):   # checks 2, 4, 6 = iterations 50, 100, 150
        X, y = [], []
        for it, tr in rows:
            if len(tr) <= at:
                continue
            cur = np.log(max(tr[at - 1][0], tr[at - 1][1]))
            prev = np.log(max(tr[at - 2][0], tr[at - 2][1]))
            slope = cur - prev                                   # per 25 iterations
            est = cur / max(-slope, 1e-3) * 25.0 if slope < 0 else 400.0
            X.append([cur, slope, min(est, 600.0), np.log(tr[at - 1][0]), np.log(tr[at - 1][1])])
            y.append(it - 25 * at)
        X, y = np.array(X), np.array(y)
        r_est = np.corrcoef(X[:, 2], y)[0, 1]
        Xa = np.c_[X, np.ones(len(X))]
        half = len(y) // 2
        w, *_ = np.linalg.lstsq(Xa[:half], y[:half], rcond=None)
        pred = Xa[half:] @ w
        r2 = 1.0 - ((pred - y[half:]) ** 2).mean() / y[half:].var()
        print(f"at iteration {25 * at}: {len(y)} still running, remaining mean {y.mean():.0f} sd {y.std():.0f}; "
              f"corr(log-linear extrapolation, remaining) {r_est:.2f}; linear model on (ratios, slope) R^2 {r2:.2f} (held out)")


if __name__ == "__main__":
    main()
