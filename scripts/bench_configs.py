"""Measures the BASELINE.json configs that are not the bench.py line (single GPU parts):
  1  single solve latency: 1000 states one at a time (seed 1001), GPU vs CPU oracle
  3  65536 states (seed 1003) on this rank's GPU(s) -- run under torchrun for 2/4/8
  4  8192 states at H = 30 (seed 1004), sharded
  5  stance-balance QP batch (seed 1005), n per GPU given on the command line
Prints one JSON object per config; results are kept under profiles/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import go1_qp_mpc_controller_b200 as pkg
from go1_qp_mpc_controller_b200.sharding import shard_range

which = sys.argv[1] if len(sys.argv) > 1 else "1"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def sync_all():
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

if which == "1":
    import oracle_binding as ob
    cfg = pkg.config_default(); eng = pkg.MpcEngine(cfg, local)
    st = pkg.generate_states(1001, 0, 1000)
    out = np.zeros(1, dtype=pkg.abi.RESULT_DTYPE)
    for i in range(20): eng.compute_grf_batch(st[i:i+1], out)
    lat = []
    for i in range(1000):
        t0 = time.perf_counter(); eng.compute_grf_batch(st[i:i+1], out); lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e3
    cl = []
    for i in range(200):
        t0 = time.perf_counter(); ob.mpc_compute_grf(cfg, st[i:i+1], threads=1); cl.append(time.perf_counter() - t0)
    cl = np.array(cl) * 1e3
    print(json.dumps({"config": 1, "what": "single Go1 MPC solve latency, H=10, host to host", "n": 1000,
                      "gpu_ms": {"mean": lat.mean(), "p50": np.percentile(lat, 50), "p99": np.percentile(lat, 99)},
                      "cpu_oracle_1thread_ms": {"mean": cl.mean(), "p50": np.percentile(cl, 50), "p99": np.percentile(cl, 99)}}))
elif which == "3":
    n = 65536; lo, hi = shard_range(n, rank, world)
    cfg = pkg.config_default(); eng = pkg.MpcEngine(cfg, local)
    st = pkg.generate_states(1003, lo, hi - lo); out = np.zeros(hi - lo, dtype=pkg.abi.RESULT_DTYPE)
    eng.compute_grf_batch(st[:1024]); eng.compute_grf_batch(st, out)
    sync_all()
    times = []
    for rep in range(5):
        sync_all(); t0 = time.perf_counter(); eng.compute_grf_batch(st, out)
        if world > 1: torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        times.append(dt)
    if rank == 0:
        print(json.dumps({"config": 3, "what": "65536 states H=10 sharded, host to host incl. H2D/D2H", "n_gpus": world,
                          "solves_per_s": n / float(np.mean(times)), "batch_ms_mean": 1e3 * float(np.mean(times)),
                          "batch_ms_p99": 1e3 * float(np.max(times)), "all_solved": bool((out["status"] == 1).all()),
                          "mean_iters": float(out["iters"].mean())}))
elif which == "4":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    lo, hi = shard_range(n, rank, world)
    cfg = pkg.config_default(); cfg.horizon = 30; eng = pkg.MpcEngine(cfg, local)
    st = pkg.generate_states(1004, lo, hi - lo); out = np.zeros(hi - lo, dtype=pkg.abi.RESULT_DTYPE)
    eng.compute_grf_batch(st[:148]); sync_all(); times = []
    for rep in range(2):
        sync_all(); t0 = time.perf_counter(); eng.compute_grf_batch(st, out); dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        times.append(dt)
    if rank == 0:
        print(json.dumps({"config": 4, "what": "long-horizon Go1 MPC H=30 (360 var), sharded, host to host", "n_gpus": world, "n": n,
                          "solves_per_s": n / float(np.mean(times)), "batch_ms_mean": 1e3 * float(np.mean(times)),
                          "all_solved": bool((out["status"] == 1).all()), "mean_iters": float(out["iters"].mean())}))
elif which == "5":
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    lo, hi = shard_range(n_total, rank, world)
    bcfg = pkg.balance_config_default(); eng = pkg.MpcEngine(bcfg, local, balance=True)
    st = pkg.generate_balance_states(1005, lo, hi - lo); out = np.zeros(hi - lo, dtype=pkg.abi.RESULT_DTYPE)
    eng.compute_grf_batch(st[:4096]); eng.compute_grf_batch(st, out)
    sync_all(); times = []
    for rep in range(5):
        sync_all(); t0 = time.perf_counter(); eng.compute_grf_batch(st, out); dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        times.append(dt)
    if rank == 0:
        st_counts = {int(k): int(v) for k, v in zip(*np.unique(out["status"], return_counts=True))}
        print(json.dumps({"config": 5, "what": "stance-balance GRF QP (12 var, 20 con), host to host", "n_gpus": world,
                          "n": n_total, "solves_per_s": n_total / float(np.mean(times)), "batch_ms_mean": 1e3 * float(np.mean(times)),
                          "status_counts_rank0": st_counts, "mean_iters": float(out["iters"].mean()), "max_iters": int(out["iters"].max())}))
if world > 1:
    dist.destroy_process_group()
