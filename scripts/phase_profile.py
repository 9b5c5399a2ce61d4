"""Per-phase cycle breakdown of admm_solve_kernel (clock64 counters of thread 0 of every CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = pkg.MpcEngine(pkg.config_default(), 0)
st = pkg.generate_states(1002, 0, n)
eng.compute_grf_batch(st)
eng.phase_cycles(True)
res = eng.compute_grf_batch(st)
pc = eng.phase_cycles(False)
tot = sum(v for k, v in pc.items() if k not in ("problems", "fine"))
npb = pc["problems"]
it = float(res["iters"].sum()); fac = float(res["rho_updates"].sum()) + n
print(f"problems {npb}  total cycles/problem {tot/npb:.0f}")
for k in ("load_scale", "factor", "iterations", "checks", "output"):
    print(f"  {k:11s} {100*pc[k]/tot:5.1f}%  {pc[k]/npb:9.0f} cyc/problem")
print(f"  per ADMM iteration {pc['iterations']/it:.0f} cyc   per factorisation {pc['factor']/fac:.0f} cyc ({pc['factor']/fac/40:.0f} per block step)   per check {pc['checks']/(it/25):.0f} cyc")
f = pc["fine"]
names = {0: "sweep: flag wait", 1: "sweep: (unused)", 2: "sweep: (unused)", 3: "sweep: owner deferred part", 4: "sweep: consumer updates",
         5: "sweep: owner panel", 8: "iter: barrier wait", 9: "iter: rhs load + 60 FMA", 10: "iter: row reduction",
         11: "iter: z/y update", 12: "iter: next rhs"}
steps = fac * 40
print("fine probes (thread 0 of each CTA), cycles per sweep step / per iteration:")
for i in (0, 1, 2, 3, 4, 5):
    print(f"  {names[i]:26s} {f[i]/steps:8.0f}")
for i in (8, 9, 10, 11, 12):
    print(f"  {names[i]:26s} {f[i]/it:8.0f}")
