"""Batches of H = 30 states through the long-horizon wrench-space engine: the ncu / WRC_PROF target.

Phase cycle counts (clock64 per phase, thread 0 of a CTA, printed for the first two problems) need a profiling build of
the library -- the same nvcc line as __graft_entry__.build() plus -DWRC_PROF:
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DWRC_PROF \
         -o go1_qp_mpc_controller_b200/libmpc_b200.so go1_qp_mpc_controller_b200/csrc/mpc_engine.cu \
         go1_qp_mpc_controller_b200/csrc/host_config.cpp -ccbin /usr/bin/g++
(rebuild with __graft_entry__.build() afterwards; each probe costs ~500 cycles, see
profiles/experiments/r02_wrench_riccati_tuning.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go1_qp_mpc_controller_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = pkg.config_default()
cfg.horizon = 30
cfg.structured_solver = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = pkg.MpcEngine(cfg, 0)
for b in range(3):
    r = e.compute_grf_batch(pkg.generate_states(1004, b * n, n))
print("ok", int((r["status"] == 1).sum()), float(r["iters"].mean()))
e.close()
