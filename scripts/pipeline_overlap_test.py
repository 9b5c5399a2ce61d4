import sys, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import go1_qp_mpc_controller_b200 as pkg
cfg = pkg.config_default()
B = 4096; steps = 20; warm = 4
rec = pkg.abi.STATE_DTYPE.itemsize
batches = [pkg.generate_states(1002, s * B, B) for s in range(steps + warm)]
pin = torch.empty((steps + warm) * B * rec, dtype=torch.uint8).pin_memory()
pn = pin.numpy().view(pkg.abi.STATE_DTYPE)
for s, b in enumerate(batches): pn[s * B:(s + 1) * B] = b
dev = pin.cuda()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
engs = [pkg.MpcEngine(cfg, 0), pkg.MpcEngine(cfg, 0)]
for e, st in zip(engs, streams): e.set_stream(st.cuda_stream)
def run(n_eng):
    for s in range(warm):
        e = engs[s % n_eng]
        e.set_states_device(dev.data_ptr() + s * B * rec, B); e.build_qp(sync=False); e.solve(sync=False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_eng)]
    e0.record(streams[0])
    for i in range(1, n_eng): streams[i].wait_event(e0)
    for k in range(steps):
        s = warm + k; e = engs[k % n_eng]
        e.set_states_device(dev.data_ptr() + s * B * rec, B); e.build_qp(sync=False); e.solve(sync=False)
    for i in range(n_eng): ends[i].record(streams[i])
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(x) for x in ends)
    return B * steps / (ms * 1e-3), ms / steps
for n_eng in (1, 2, 1, 2):
    v, ms = run(n_eng)
    print(f"engines {n_eng}: {v:,.0f} solves/s  {ms:.3f} ms/step")
# e2e with two host threads
out = [np.empty(B, dtype=pkg.abi.RESULT_DTYPE) for _ in range(2)]
def worker(i, n_eng, lst):
    for k in range(i, steps, n_eng):
        s = warm + k
        engs[i].compute_grf_batch(pn[s * B:(s + 1) * B], out[i])
for n_eng in (1, 2, 1, 2):
    for i in range(n_eng): engs[i].compute_grf_batch(pn[:B], out[i])
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i, n_eng, None)) for i in range(n_eng)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"e2e threads {n_eng}: {B * steps / dt:,.0f} solves/s")
