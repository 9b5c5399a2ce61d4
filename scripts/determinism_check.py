import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
cfg=pkg.config_default(); e=pkg.MpcEngine(cfg,0)
st=pkg.generate_states(1002,0,1024)
runs=[e.compute_grf_batch(st)["iters"].copy() for _ in range(int(sys.argv[1]))]
ref=np.array(runs); mode=np.median(ref,axis=0)
bad=[(r, list(np.nonzero(ref[r]!=mode)[0][:8])) for r in range(len(runs)) if (ref[r]!=mode).any()]
print("runs differing from the per-problem median:", len(bad), bad[:4])
