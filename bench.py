#!/usr/bin/env python
"""bench.py -- batched MPC QP solves/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (QP build + ADMM solve + result write) over one batch
of synthetic robot states.  Workload = BASELINE.json configs[1]: 4096 synthetic Go1 states,
horizon 10, gazebo_a1_mpc.yaml weights, eps_abs = eps_rel = 1e-5, per GPU (weak scaling:
every rank gets its own 4096-state shard of the counter-based stream; no data-path
collective, one final gather).  One JSON line on rank 0:

  value      whole-job solves/s with the state records already resident in HBM
  e2e        the same metric through the reference-facing C ABI call with HOST buffers
             (mpc_compute_grf_batch: H2D + build + solve + D2H inside the timed region)
  roofline   dominant kernel (admm_solve_kernel): algorithmic flops / CUDA-event duration
             against this box's measured FP64 FMA peak
  cpu_baseline  the oracle (CPU port of the reference path) timed on the host cores

--impl reference times the reference's own CPU algorithm (the oracle port: the real
Eigen/OsqpEigen/OSQP stack cannot be built here) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4096
SEED = 1002
WORKLOAD = "batched Go1 MPC H=10, 4096 synthetic robot states per GPU (BASELINE configs[1])"
# SURVEY.md 8d algorithmic-flop model, dense formulation the reference executes, H = 10
F_BUILD = 4.00e6
F_FACTOR = 0.576e6
F_ITER = 33.0e3
# FP64 FMA peak of this pool's B200, measured with scripts/fp64_bench.cu
# (profiles/r01_fp64_peak.txt): 17.07 T DFMA/s = 34.1 TFLOP/s
FP64_PEAK_TFLOPS = 34.1
# dram__bytes_read.sum + dram__bytes_write.sum of admm_solve_kernel per solve, from the committed
# ncu --set full capture (profiles/r01_v13_ncu_summary.txt: 37.30 MB read, 0 written, 296 solves):
# the padded f64 Hessian (122,880 B) + q, l, u, state.  P is an intermediate of the path, not
# algorithmic input, hence the much smaller hbm_algorithmic_bytes_per_solve.
NCU_DRAM_BYTES_PER_SOLVE = 37.295872e6 / 296


def measured_hbm_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 7700.0, "fallback (B200_PROFILING.md nominal)"


def host_cores():
    """Host cores this process may use (all of them: the CPU arm is one problem per thread)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region: NVML every 10 ms when the bindings load
    (the timed region is a fraction of a second), else the nvidia-smi query line every 0.2 s."""

    # nvmlClocksEventReason* bits
    _BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = None
        self._halt = threading.Event()
        self._go = threading.Event()  # samples are kept only between begin() and stop()

    def begin(self):
        self._go.set()

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        self.source = "nvml"
        while not self._halt.is_set():
            if self._go.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for name, bit in self._BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            self._halt.wait(0.01 if self._go.is_set() else 0.001)
        nv.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self._halt.is_set():
            if not self._go.is_set():
                self._halt.wait(0.001)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    r = [c.strip() for c in out.split(",")]
                    if r[0].replace(".", "").isdigit():
                        self.sm.append(float(r[0]))
                    if r[1].replace(".", "").isdigit():
                        self.mx.append(float(r[1]))
                    for nme, v in zip(names, r[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(nme)
            except Exception:
                pass
            self._halt.wait(0.2)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            if not self._halt.is_set():
                self._run_smi()

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def run_reference(args):
    """The reference's CPU algorithm (oracle port) on the host cores, same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import go1_qp_mpc_controller_b200 as pkg
    import oracle_binding as ob
    cfg = pkg.config_default()
    threads = host_cores()  # torchrun exports OMP_NUM_THREADS=1; num_threads() overrides it
    sample = 1024  # bounded sample of the 4096-state batch per step
    times = []
    for step in range(args.warmup + args.steps):
        states = pkg.generate_states(SEED, step * BATCH, sample)
        t0 = time.perf_counter()
        res = ob.mpc_compute_grf(cfg, states, threads=threads)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": "batched MPC QP solves/sec (H=10)", "value": value,
        "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "horizon": 10, "eps_abs": 1e-5, "eps_rel": 1e-5,
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} states of the 4096-state batch per step, OpenMP one problem per thread"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_iters": float(res["iters"].mean()),
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    import go1_qp_mpc_controller_b200 as pkg
    from go1_qp_mpc_controller_b200.sharding import gather_results

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cfg = pkg.config_default()
    eng = pkg.MpcEngine(cfg, local_rank)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    nsteps = args.warmup + args.steps
    rec = pkg.abi.STATE_DTYPE.itemsize
    # fresh batch per step and per rank: global state index = (step * world + rank) * BATCH + i
    host_batches = [pkg.generate_states(SEED, (s * world + rank) * BATCH, BATCH) for s in range(nsteps)]
    pinned_in = torch.empty(nsteps * BATCH * rec, dtype=torch.uint8).pin_memory()
    pin_np = pinned_in.numpy().view(pkg.abi.STATE_DTYPE)
    for s in range(nsteps):
        pin_np[s * BATCH:(s + 1) * BATCH] = host_batches[s]
    dev_in = pinned_in.cuda()
    pinned_out = torch.empty(BATCH * pkg.abi.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out_np = pinned_out.numpy().view(pkg.abi.RESULT_DTYPE)

    def step_device(s):
        eng.set_states_device(dev_in.data_ptr() + s * BATCH * rec, BATCH)
        eng.build_qp(sync=False)
        eng.solve(sync=False)

    # ---------------- value: inputs resident in HBM, device-timed ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # NVML comes up during the warm-up; samples are kept from begin() on
    with torch.cuda.stream(stream):
        for s in range(args.warmup):
            step_device(s)
    barrier()
    sampler.begin()
    launches0 = eng.kernel_launches()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for k in range(args.steps):
            s = args.warmup + k
            eng.set_states_device(dev_in.data_ptr() + s * BATCH * rec, BATCH)
            ev[k][0].record()
            eng.build_qp(sync=False)
            ev[k][1].record()
            eng.solve(sync=False)
            ev[k][2].record()
        e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    build_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]))
    solve_ms = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)]))
    last = eng.get_results()
    ok = bool((last["status"] == 1).all())
    mean_iters = sum_over_ranks(float(last["iters"].mean())) / world
    mean_fac = 1.0 + sum_over_ranks(float(last["rho_updates"].mean())) / world
    value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---------------- e2e: host buffers through the C ABI, H2D and D2H inside ----------------
    lat = []
    for s in range(args.warmup):
        eng.compute_grf_batch(pin_np[s * BATCH:(s + 1) * BATCH], out_np)
    barrier()
    t_start = time.perf_counter()
    for k in range(args.steps):
        s = args.warmup + k
        t0 = time.perf_counter()
        eng.compute_grf_batch(pin_np[s * BATCH:(s + 1) * BATCH], out_np)
        lat.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t_start)
    barrier()
    e2e_value = world * BATCH * args.steps / e2e_s
    last_out = out_np.copy()  # results of the last timed batch: the CPU baseline checks parity on these
    # the single exchange of the path: final gather of the 64 B records (outside the step loop)
    if world > 1:
        g0 = time.perf_counter()
        full = gather_results(out_np.copy(), BATCH * world)
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - g0)
        assert rank != 0 or len(full) == BATCH * world
    else:
        gather_ms = 0.0

    # ---------------- extra: the widened path, warm-started streaming ticks ----------------
    # one persistent solver per robot slot, consecutive control ticks of the same robots, host
    # buffers in and out every tick; every rank streams its own BATCH robots
    stream_warm = None
    try:
        ticks = 12
        sbuf = torch.empty(ticks * BATCH * rec, dtype=torch.uint8).pin_memory()
        s_np = sbuf.numpy().view(pkg.abi.STATE_DTYPE)
        for t in range(ticks):
            s_np[t * BATCH:(t + 1) * BATCH] = pkg.generate_stream_states(SEED, rank * BATCH, BATCH, 40 + t)
        eng.stream_reset()
        eng.stream_step(s_np[:BATCH], out_np)          # tick 0 is the cold initSolver tick
        eng.stream_step(s_np[BATCH:2 * BATCH], out_np)
        barrier()
        t0 = time.perf_counter()
        it_sum = 0.0
        for t in range(2, ticks):
            eng.stream_step(s_np[t * BATCH:(t + 1) * BATCH], out_np)
            it_sum += float(out_np["iters"].mean())
        torch.cuda.synchronize()
        sdt = max_over_ranks(time.perf_counter() - t0)
        stream_warm = {"metric": "warm-started MPC robot-ticks/sec (H=10)", "value": world * BATCH * (ticks - 2) / sdt,
                  "unit": "robot-ticks/s", "ticks_timed": ticks - 2, "ms_per_tick": 1e3 * sdt / (ticks - 2),
                  "mean_iters": it_sum / (ticks - 2), "all_solved": bool((out_np["status"] == 1).all())}
    except Exception as ex:  # the headline line must not depend on the extra
        stream_warm = {"error": str(ex)}
    barrier()

    # ---------------- extra: BASELINE configs[3], long horizon H = 30 (360 variables) ----------------
    # Riccati-structured solver, 2048 synthetic states per GPU (seed 1004), host buffers in and out
    long_horizon = None
    try:
        cfg30 = pkg.config_default()
        cfg30.horizon = 30
        eng30 = pkg.MpcEngine(cfg30, local_rank)
        n30 = 2048
        st30 = pkg.generate_states(1004, rank * n30, n30)
        eng30.compute_grf_batch(st30)  # full-size warm-up: the engine sizes its buffers on first use
        barrier()
        t0 = time.perf_counter()
        reps30 = 3
        for _ in range(reps30):
            out30 = eng30.compute_grf_batch(st30)
        torch.cuda.synchronize()
        dt30 = max_over_ranks(time.perf_counter() - t0)
        long_horizon = {"metric": "batched MPC QP solves/sec (H=30)", "value": world * n30 * reps30 / dt30,
                        "unit": "solves/s", "states_per_gpu": n30, "ms_per_batch": 1e3 * dt30 / reps30,
                        "mean_iters": float(out30["iters"].mean()),
                        "all_solved": bool((out30["status"] == 1).all())}
        eng30.close()
    except Exception as ex:  # the headline line must not depend on the extra
        long_horizon = {"error": str(ex)}
    barrier()

    # ---------------- CPU baseline: the oracle on the host cores (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        threads = host_cores()
        sample = 2048
        t0 = time.perf_counter()
        ref = ob.mpc_compute_grf(cfg, host_batches[nsteps - 1][:sample], threads=threads)
        cdt = time.perf_counter() - t0
        den = np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
        rel = np.linalg.norm(last_out["grf"][:sample].astype(np.float64) - ref["grf"], axis=1) / den
        cpu = {"value": sample / cdt, "unit": "solves/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} states of the last timed batch, OpenMP one problem per thread, fp64",
               "parity_max_rel_grf_err": float(rel.max()),
               "parity_same_iters": float((ref["iters"] == last_out["iters"][:sample]).mean())}

    if rank == 0:
        flops_solve = mean_fac * F_FACTOR + mean_iters * F_ITER          # admm_solve_kernel, per solve
        achieved = BATCH * flops_solve / (solve_ms * 1e-3) / 1e12
        line = {
            "metric": "batched MPC QP solves/sec (H=10)", "value": value, "unit": "solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "horizon": 10, "states_per_gpu": BATCH,
                       "weights": "config/gazebo_a1_mpc.yaml", "eps_abs": 1e-5, "eps_rel": 1e-5,
                       "max_iter": 4000, "adaptive_rho_interval": 50, "cold_start": True,
                       "l2": "fresh state batch per step; per-step working set (P, 236 MB) exceeds the 126 MB L2",
                       "parallelism": f"shard{world}"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": BATCH * rec,
                    "d2h_bytes_per_step": BATCH * pkg.abi.RESULT_DTYPE.itemsize,
                    "p50_batch_ms": 1e3 * float(np.percentile(lat, 50)),
                    "p99_batch_ms": 1e3 * float(np.percentile(lat, 99))},
            "gpu_launches": int(launches),
            "kernels": {"qp_build_kernel_ms": build_ms, "admm_solve_kernel_ms": solve_ms,
                        "final_gather_ms": gather_ms},
            "roofline": {"bound": "fp64-fma (compute/latency; neither hbm nor tensor, SURVEY.md 8d)",
                         "kernel": "admm_solve_kernel", "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                         "traffic": NCU_DRAM_BYTES_PER_SOLVE * BATCH,
                         "traffic_source": "ncu --set full, profiles/r01_v19_ncu_summary.txt, scaled to this launch's solves",
                         "hbm": {"achieved": NCU_DRAM_BYTES_PER_SOLVE * BATCH / (solve_ms * 1e-3) / 1e9,
                                 "peak": measured_hbm_peak_gbs()[0], "unit": "GB/s",
                                 "frac": NCU_DRAM_BYTES_PER_SOLVE * BATCH / (solve_ms * 1e-3) / 1e9 / measured_hbm_peak_gbs()[0],
                                 "peak_source": measured_hbm_peak_gbs()[1]},
                         "peak_source": "measured on this pool: scripts/fp64_bench.cu, profiles/r01_fp64_peak.txt",
                         "algorithmic_flops_per_solve": flops_solve,
                         "hbm_algorithmic_bytes_per_solve": 256},
            "solver": {"mean_iters": mean_iters, "mean_factorisations": mean_fac, "all_solved": ok},
            "stream_warm": stream_warm,
            "long_horizon_h30": long_horizon,
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
