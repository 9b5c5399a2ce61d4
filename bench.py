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
             (mpc_compute_grf_batch: H2D + fused build/solve + D2H inside the timed region; at N > 1
             also the gather of every rank's results into one host array on rank 0, every step)
  roofline   the one kernel of the step (wrench_tile_kernel): algorithmic flops / CUDA-event
             duration against this GPU's FP64 FMA rate measured in the same run
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
# FP64 FMA peak: measured live by mpc_measure_fp64_peak in every run; this constant (round 1,
# scripts/fp64_bench.cu, profiles/r01_fp64_peak.txt) is only the fallback if the probe fails
FP64_PEAK_TFLOPS = 34.1
# dram__bytes_read.sum + dram__bytes_write.sum of wrench_tile_kernel per solve, from the committed
# ncu --set full capture (profiles/r02_wrench_tile_ncu_summary.txt: 0.920 MB read + 1.487 MB written by
# one launch of 4096 solves): the 192 B record in, the 64 B result and the 480 B primal solution out, plus
# whatever of the kernel's local-memory (spill) lines L2 wrote back -- between 0.7 and 7.8 MB per launch across
# the round's captures.  (Round 1's two-kernel path moved 126 KB per solve for the f64 Hessian.)
NCU_DRAM_BYTES_PER_SOLVE = (919.552e3 + 1486.848e3) / 4096


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


# --------------------------------------------------------------------------------------------
# host-only access to configuration defaults and the synthetic generator (libmpc_hostgen.so is
# host_config.cpp compiled alone): the reference arm never maps the product library
# --------------------------------------------------------------------------------------------
class HostGen:
    def __init__(self):
        import ctypes as C
        import __graft_entry__ as g
        from go1_qp_mpc_controller_b200 import abi     # ctypes / numpy record layouts only, loads nothing
        if not os.path.exists(g.HOSTGEN):
            g.build()
        self.C, self.abi = C, abi
        self.lib = C.CDLL(g.HOSTGEN)
        self.lib.mpc_generate_states.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p]

    def config_default(self):
        cfg = self.abi.MpcConfig()
        self.lib.mpc_config_default(self.C.byref(cfg))
        return cfg

    def generate_states(self, seed, first, n):
        out = np.zeros(n, dtype=self.abi.STATE_DTYPE)
        rc = self.lib.mpc_generate_states(seed, first, n, out.ctypes.data_as(self.C.c_void_p))
        assert rc == 0
        return out


def run_reference(args):
    """The reference's CPU algorithm (oracle port; the real Eigen/OsqpEigen/OSQP stack cannot be built
    here) on ALL host cores, on exactly the batches rank 0 of the GPU arm solves: the full
    4096-state batch of every step, same stream indices, same settings."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    hg = HostGen()
    cfg = hg.config_default()
    flags = ob.use_native()
    threads = host_cores()  # torchrun exports OMP_NUM_THREADS=1; num_threads() overrides it
    world = max(1, args.gpus)
    times = []
    for step in range(args.warmup + args.steps):
        states = hg.generate_states(SEED, (step * world + 0) * BATCH, BATCH)   # rank 0's batch of this step
        t0 = time.perf_counter()
        res = ob.mpc_compute_grf(cfg, states, threads=threads)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = BATCH * args.steps / total
    line = {
        "impl": "reference", "metric": "batched MPC QP solves/sec (H=10)", "value": value,
        "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "horizon": 10, "states_per_gpu": BATCH,
                   "weights": "config/gazebo_a1_mpc.yaml", "eps_abs": 1e-5, "eps_rel": 1e-5,
                   "max_iter": 4000, "adaptive_rho_interval": 50, "cold_start": True,
                   "sample_per_step": BATCH, "stream_indices": "rank 0's batches of the GPU arm"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": f"all {BATCH} states of every step, OpenMP one problem per thread, fp64",
                         "compiler_flags": flags,
                         "p50_batch_ms": 1e3 * float(np.percentile(times, 50)),
                         "p99_batch_ms": 1e3 * float(np.percentile(times, 99))},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_iters": float(res["iters"].mean()),
    }
    print(json.dumps(line), flush=True)
    return 0


class DevBuf:
    """A device pointer as a __cuda_array_interface__ object, so that torch can view the engine's result
    buffer without a copy (the device-side gather reads it in place)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import go1_qp_mpc_controller_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # host-side barrier (gloo): a rank waiting in an NCCL barrier spins a kernel on its GPU, which would
    # time-slice against rank 0's process when that one drives every GPU itself (fleet extra)
    host_group = dist.new_group(backend="gloo") if world > 1 else None

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
    rsz = pkg.abi.RESULT_DTYPE.itemsize
    # fresh batch per step and per rank: global state index = (step * world + rank) * BATCH + i
    host_batches = [pkg.generate_states(SEED, (s * world + rank) * BATCH, BATCH) for s in range(nsteps)]
    pinned_in = torch.empty(nsteps * BATCH * rec, dtype=torch.uint8).pin_memory()
    pin_np = pinned_in.numpy().view(pkg.abi.STATE_DTYPE)
    for s in range(nsteps):
        pin_np[s * BATCH:(s + 1) * BATCH] = host_batches[s]
    dev_in = pinned_in.cuda()
    pinned_out = torch.empty(BATCH * rsz, dtype=torch.uint8).pin_memory()
    out_np = pinned_out.numpy().view(pkg.abi.RESULT_DTYPE)

    def step_device(s):
        eng.set_states_device(dev_in.data_ptr() + s * BATCH * rec, BATCH)
        eng.build_qp(sync=False)     # no launch: the wrench-space engine builds inside the solve kernel
        eng.solve(sync=False)

    # ---------------- FP64 roofline denominator, measured now, with the clocks it ran at ----------------
    peak_sampler = ClockSampler(local_rank)
    fp64_peak = None
    if rank == 0:
        peak_sampler.start()
        pkg.measure_fp64_peak(local_rank, 2.0)       # warm-up; NVML comes up meanwhile
        peak_sampler.begin()
        fp64_peak = pkg.measure_fp64_peak(local_rank, 60.0)
        peak_clocks = peak_sampler.stop()
    barrier()

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
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for k in range(args.steps):
            s = args.warmup + k
            eng.set_states_device(dev_in.data_ptr() + s * BATCH * rec, BATCH)
            eng.build_qp(sync=False)
            ev[k][0].record()
            eng.solve(sync=False)
            ev[k][1].record()
        e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    solve_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]))
    last = eng.get_results()
    ok = bool((last["status"] == 1).all())
    mean_iters = sum_over_ranks(float(last["iters"].mean())) / world
    mean_fac = 1.0 + sum_over_ranks(float(last["rho_updates"].mean())) / world
    value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---------------- e2e: host buffers through the C ABI, H2D and D2H inside; at N > 1 the fleet's ----------------
    # results are ALSO gathered into one host array on rank 0 inside the timed region, every step:
    # NCCL gather of the 64 B records straight from each engine's device result buffer, then one D2H copy
    gather_out = gather_list = None
    if world > 1:
        res_view = None
        if rank == 0:
            gather_dev = torch.empty(world * BATCH * rsz, dtype=torch.uint8, device="cuda")
            gather_list = list(gather_dev.chunk(world))          # views: the gather lands contiguously
            gather_out = torch.empty(world * BATCH * rsz, dtype=torch.uint8).pin_memory()

    def e2e_step(s):
        if world == 1:
            eng.compute_grf_batch(pin_np[s * BATCH:(s + 1) * BATCH], out_np)
            return
        # N > 1: everything of a step is ordered on the engine's stream and the host waits ONCE -- H2D of the host
        # states (mpc_load_states), the solve, the NCCL gather straight from the engine's device result buffer, and the
        # device-to-host copy: all ranks' records into one pinned array on rank 0, the own records elsewhere
        with torch.cuda.stream(stream):
            eng.load_states(pin_np[s * BATCH:(s + 1) * BATCH])
            eng.build_qp(sync=False)
            eng.solve(sync=False)
            view = torch.as_tensor(DevBuf(eng.results_device_ptr(), BATCH * rsz), device="cuda")
            dist.gather(view, gather_list, dst=0)
            if rank == 0:
                gather_out.copy_(gather_dev, non_blocking=True)
            else:
                pinned_out.copy_(view, non_blocking=True)
        stream.synchronize()

    lat = []
    for s in range(args.warmup):
        e2e_step(s)
    barrier()
    t_start = time.perf_counter()
    for k in range(args.steps):
        t0 = time.perf_counter()
        e2e_step(args.warmup + k)
        lat.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t_start)
    barrier()
    e2e_value = world * BATCH * args.steps / e2e_s
    gather_ok = None
    if world > 1 and rank == 0:
        full = gather_out.numpy().view(pkg.abi.RESULT_DTYPE)
        out_np[:] = full[:BATCH]              # rank 0's own shard is the first chunk of the gathered array
        own = eng.get_results()               # the same records read back from the engine: the gather moved them intact
        gather_ok = bool(np.array_equal(full[:BATCH]["grf"], own["grf"]) and (full["status"] == 1).all())
    last_out = out_np.copy()  # results of the last timed batch: the CPU baseline checks parity on these

    extras = {}

    def extra(name, fn):
        try:
            extras[name] = fn()
        except Exception as ex:  # the headline line must not depend on an extra
            extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
        barrier()

    # ---------------- extra: the widened path, warm-started streaming ticks ----------------
    def x_stream():
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
        return {"metric": "warm-started MPC robot-ticks/sec (H=10)", "value": world * BATCH * (ticks - 2) / sdt,
                "unit": "robot-ticks/s", "ticks_timed": ticks - 2, "ms_per_tick": 1e3 * sdt / (ticks - 2),
                "mean_iters": it_sum / (ticks - 2), "all_solved": bool((out_np["status"] == 1).all())}
    extra("stream_warm", x_stream)

    # ---------------- extra: BASELINE configs[2], 65536 states over the N GPUs (strong scaling) ----------------
    def x_strong():
        total = 65536
        lo, hi = pkg.fleet_shard_range(total, world, rank)
        st = pkg.generate_states(1003, lo, hi - lo)
        buf = torch.empty((hi - lo) * rec, dtype=torch.uint8).pin_memory()
        b_np = buf.numpy().view(pkg.abi.STATE_DTYPE)
        b_np[:] = st
        o = torch.empty((hi - lo) * rsz, dtype=torch.uint8).pin_memory().numpy().view(pkg.abi.RESULT_DTYPE)
        eng.compute_grf_batch(b_np, o)
        barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.compute_grf_batch(b_np, o)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return {"metric": "batched MPC QP solves/sec (H=10), 65536 states sharded over the GPUs", "scaling": "strong",
                "value": total * reps / dt, "unit": "solves/s", "states_total": total, "states_per_gpu": hi - lo,
                "ms_per_batch": 1e3 * dt / reps, "all_solved": bool((o["status"] == 1).all())}
    extra("config3_65536_sharded", x_strong)

    # ---------------- extra: BASELINE configs[3], long horizon H = 30 (360 variables) ----------------
    def x_h30():
        n30 = max(8192 // world, 1024)
        st30 = pkg.generate_states(1004, rank * n30, n30)

        def run30(solver, reps30):
            cfg30 = pkg.config_default()
            cfg30.horizon = 30
            cfg30.structured_solver = solver
            eng30 = pkg.MpcEngine(cfg30, local_rank)
            eng30.compute_grf_batch(st30)  # full-size warm-up: the engine sizes its buffers on first use
            l0 = eng30.kernel_launches()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps30):
                out = eng30.compute_grf_batch(st30)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            launches = (eng30.kernel_launches() - l0) // reps30
            eng30.close()
            return world * n30 * reps30 / dt, 1e3 * dt / reps30, out, launches
        def warm30():
            # the controller's streaming use at the long horizon: one persistent warm-started solver per robot
            cfg30 = pkg.config_default()
            cfg30.horizon = 30
            eng30 = pkg.MpcEngine(cfg30, local_rank)
            nw, ticks = 2048, 8
            sts = [pkg.generate_stream_states(SEED, rank * nw, nw, 40 + t) for t in range(ticks)]
            o = np.zeros(nw, dtype=pkg.abi.RESULT_DTYPE)
            eng30.stream_reset()
            eng30.stream_step(sts[0], o)
            eng30.stream_step(sts[1], o)
            barrier()
            t0 = time.perf_counter()
            its = 0.0
            for t in range(2, ticks):
                eng30.stream_step(sts[t], o)
                its += float(o["iters"].mean())
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            ok = bool((o["status"] == 1).all())
            eng30.close()
            return {"value": world * nw * (ticks - 2) / dt, "unit": "robot-ticks/s", "robots_per_gpu": nw,
                    "ms_per_tick": 1e3 * dt / (ticks - 2), "mean_iters": its / (ticks - 2), "all_solved": ok}
        v3, ms3, out30, l3 = run30(0, 3)      # default at H = 30: wrench_riccati_kernel (fused build + six-input Riccati ADMM)
        v1, ms1, out1, l1 = run30(1, 1)       # round 1's engine: dense build + riccati_solve_kernel
        return {"metric": "batched MPC QP solves/sec (H=30)", "value": v3, "unit": "solves/s", "states_per_gpu": n30,
                "ms_per_batch": ms3, "kernel": "wrench_riccati_kernel<30>", "launches_per_batch": int(l3),
                "mean_iters": float(out30["iters"].mean()), "all_solved": bool((out30["status"] == 1).all()),
                "riccati_engine": {"value": v1, "ms_per_batch": ms1, "launches_per_batch": int(l1),
                                   "same_iterations": bool(np.array_equal(out1["iters"], out30["iters"]))},
                "stream_warm": warm30()}
    extra("long_horizon_h30", x_h30)

    # ---------------- extra: BASELINE configs[4], stance-balance QP, 1 M problems over 8 GPUs ----------------
    def x_balance():
        bcfg = pkg.balance_config_default()
        be = pkg.MpcEngine(bcfg, local_rank, balance=True)
        nb = 125000                                       # the per-GPU share of 1 M problems on 8 GPUs
        stb0 = pkg.generate_balance_states(1005, rank * nb, nb)
        # page-locked host buffers, as in the headline e2e (a pageable array costs a staged copy of 32 MB per batch)
        pin_i = torch.empty(nb * stb0.dtype.itemsize, dtype=torch.uint8).pin_memory()
        stb = pin_i.numpy().view(stb0.dtype)
        stb[:] = stb0
        ob_ = torch.empty(nb * rsz, dtype=torch.uint8).pin_memory().numpy().view(pkg.abi.RESULT_DTYPE)
        be.compute_grf_batch(stb, ob_)
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            be.compute_grf_batch(stb, ob_)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        # SURVEY.md 8d flop model of the 12-variable QP: 2.3 k (build + factor) + 0.6 k per iteration
        it = float(ob_["iters"].mean())
        fl = 2.3e3 + 0.6e3 * it
        r = {"metric": "stance-balance QP solves/sec (12 var / 20 con)", "value": world * nb * reps / dt,
             "unit": "solves/s", "problems_per_gpu": nb, "ms_per_batch": 1e3 * dt / reps, "mean_iters": it,
             "max_iters": int(ob_["iters"].max()), "all_solved": bool((ob_["status"] == 1).all()),
             "kernel": "balance_qp_leg_kernel (four lanes per problem, eight problems per warp)",
             "host_buffers": "page-locked",
             "roofline": {"bound": "fp64-fma", "achieved": nb * reps * fl / dt / 1e12, "unit": "TFLOP/s",
                          "algorithmic_flops_per_solve": fl, "note": "host to host, per GPU"}}
        be.close()
        return r
    extra("balance_qp", x_balance)

    # ---------------- extra: BASELINE configs[0], single-solve latency, GPU and CPU (N = 1, rank 0) ----------------
    def x_latency():
        if world > 1:
            return {"skipped": "replicas only at N > 1 (DESIGN.md 6)"}
        one = pkg.generate_states(1001, 0, 300)
        o1 = np.zeros(1, dtype=pkg.abi.RESULT_DTYPE)
        for i in range(20):
            eng.compute_grf_batch(one[i:i + 1], o1)
        g = []
        for i in range(200):
            t0 = time.perf_counter()
            eng.compute_grf_batch(one[i:i + 1], o1)
            g.append(time.perf_counter() - t0)
        r = {"metric": "single Go1 MPC solve latency (H=10), host to host", "unit": "ms", "solves": 200,
             "gpu_p50_ms": 1e3 * float(np.percentile(g, 50)), "gpu_p99_ms": 1e3 * float(np.percentile(g, 99))}
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_binding as ob
            ob.use_native()
            c = []
            for i in range(100):
                t0 = time.perf_counter()
                ob.mpc_compute_grf(cfg, one[i:i + 1], threads=1)
                c.append(time.perf_counter() - t0)
            r.update({"cpu_p50_ms": 1e3 * float(np.percentile(c, 50)), "cpu_p99_ms": 1e3 * float(np.percentile(c, 99)),
                      "cpu_kind": "port, one thread"})
        return r
    extra("latency_single", x_latency)

    # ---------------- extra: every GPU of the box through ONE C-ABI call from ONE process (rank 0) ----------------
    def x_fleet():
        ndev = torch.cuda.device_count() if world == 1 else world
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=host_group)      # every GPU idle from here on
        if rank != 0:
            dist.barrier(group=host_group)      # CPU wait while rank 0's process uses all the GPUs
            return None
        fl = pkg.MpcFleet(cfg, list(range(ndev)))
        n = BATCH * ndev
        stf = pkg.generate_states(SEED, 0, n)
        fin = torch.empty(n * rec, dtype=torch.uint8).pin_memory()
        f_np = fin.numpy().view(pkg.abi.STATE_DTYPE)
        f_np[:] = stf
        fo = torch.empty(n * rsz, dtype=torch.uint8).pin_memory().numpy().view(pkg.abi.RESULT_DTYPE)
        fl.compute_grf_batch(f_np, fo)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            fl.compute_grf_batch(f_np, fo)
        dt = time.perf_counter() - t0
        r = {"metric": "batched MPC QP solves/sec (H=10) through mpc_fleet_compute_grf_batch, one process",
             "value": n * reps / dt, "unit": "solves/s", "devices": ndev, "states_total": n,
             "ms_per_batch": 1e3 * dt / reps, "all_solved": bool((fo["status"] == 1).all()),
             "note": "host array in, ONE host array out; per-GPU cudaMemcpyAsync is the gather, no NCCL"}
        fl.close()
        if world > 1:
            dist.barrier(group=host_group)
        return r
    extra("fleet_c_abi", x_fleet)

    # ---------------- CPU baseline: the oracle on the host cores (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        flags = ob.use_native()
        threads = host_cores()
        ct = []
        for b in range(3):                                # three whole batches: a CPU batch latency beside the GPU's
            t0 = time.perf_counter()
            ref = ob.mpc_compute_grf(cfg, host_batches[nsteps - 1 - b], threads=threads)
            ct.append(time.perf_counter() - t0)
            if b == 0:
                ref_last = ref
        den = np.maximum(np.linalg.norm(ref_last["grf"], axis=1), 1.0)
        rel = np.linalg.norm(last_out["grf"].astype(np.float64) - ref_last["grf"], axis=1) / den
        cpu = {"value": 3 * BATCH / float(np.sum(ct)), "unit": "solves/s", "cores": threads, "kind": "port",
               "sample": f"the last three timed batches ({3 * BATCH} states), OpenMP one problem per thread, fp64",
               "compiler_flags": flags,
               "p50_batch_ms": 1e3 * float(np.percentile(ct, 50)), "p99_batch_ms": 1e3 * float(np.max(ct)),
               "parity_max_rel_grf_err": float(rel.max()),
               "parity_same_iters": float((ref_last["iters"] == last_out["iters"]).mean())}

    if rank == 0:
        flops_solve = F_BUILD + mean_fac * F_FACTOR + mean_iters * F_ITER   # build is inside the kernel now
        achieved = BATCH * flops_solve / (solve_ms * 1e-3) / 1e12
        peak = fp64_peak if fp64_peak else FP64_PEAK_TFLOPS
        hbm_peak, hbm_src = measured_hbm_peak_gbs()
        traffic = NCU_DRAM_BYTES_PER_SOLVE * BATCH
        line = {
            "metric": "batched MPC QP solves/sec (H=10)", "value": value, "unit": "solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "horizon": 10, "states_per_gpu": BATCH,
                       "weights": "config/gazebo_a1_mpc.yaml", "eps_abs": 1e-5, "eps_rel": 1e-5,
                       "max_iter": 4000, "adaptive_rho_interval": 50, "cold_start": True,
                       "l2": "fresh state batch per step; the path keeps no Hessian in memory (0.4 KB of HBM traffic "
                             "per solve), so there is nothing for the L2 to retain between steps",
                       "parallelism": f"shard{world}"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": BATCH * rec,
                    "d2h_bytes_per_step": BATCH * rsz * world,   # rank 0 (N > 1: all ranks' records in one copy)
                    "p50_batch_ms": 1e3 * float(np.percentile(lat, 50)),
                    "p99_batch_ms": 1e3 * float(np.percentile(lat, 99)),
                    "gather": (None if world == 1 else
                               {"inside_timed_region": True, "how": "NCCL gather of the 64 B records from every engine's device "
                                "result buffer to rank 0, then one D2H copy into one pinned host array, every step; "
                                "stream-ordered behind the solve, one host wait per step",
                                "verified": gather_ok})},
            "gpu_launches": int(launches),
            "kernels": {"wrench_tile_kernel_ms": solve_ms, "launches_per_step": launches / max(1, args.steps)},
            "roofline": {"bound": "fp64-fma (compute/latency; neither hbm nor tensor, SURVEY.md 8d)",
                         "kernel": "wrench_tile_kernel", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": ("mpc_measure_fp64_peak on this GPU in this run (register-only DFMA kernel, all SMs)"
                                         if fp64_peak else "profiles/r01_fp64_peak.txt (live probe failed)"),
                         "peak_clocks": (peak_clocks if fp64_peak else None),
                         "traffic": traffic,
                         "traffic_source": "ncu --set full, profiles/r02_wrench_tile_ncu_summary.txt (dram read + write of one "
                                           "4096-solve launch), scaled to this launch's solves",
                         "hbm": {"achieved": traffic / (solve_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": traffic / (solve_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                         "algorithmic_flops_per_solve": flops_solve,
                         "algorithmic_flop_model": "SURVEY.md 8d dense formulation: 4.00 M (build) + n_fac 0.576 M + n_iter 33 k; "
                                                   "the kernel executes about a third of that (rank-60 structure)",
                         "hbm_algorithmic_bytes_per_solve": 256},
            "solver": {"mean_iters": mean_iters, "mean_factorisations": mean_fac, "all_solved": ok},
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        line.update(extras)
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
