"""Multi-GPU inside the C ABI (mpc_fleet_*, SURVEY.md 8b "device list" / 8e): contiguous shards, one
stream and one pinned staging pair per GPU, results into ONE caller-owned host array.  A device may be
listed twice, so the sharding logic is exercised on a one-GPU box; the 2/4/8-GPU numbers are bench.py's."""
import numpy as np
import pytest


def test_shard_range_is_the_contiguous_ceil_partition(pkg):
    from go1_qp_mpc_controller_b200.sharding import shard_range
    for n in (0, 1, 7, 4096, 65536, 1000003):
        for g in (1, 2, 3, 4, 8):
            prev = 0
            for i in range(g):
                b, e = pkg.fleet_shard_range(n, g, i)
                assert (b, e) == shard_range(n, i, g)
                assert b == prev and e >= b and (e - b) - n // g in (0, 1)
                prev = e
            assert prev == n
    with pytest.raises(pkg.MpcError):
        pkg.fleet_shard_range(10, 0, 0)


def test_fleet_fails_loudly_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.MpcError) as ei:
        pkg.MpcFleet(pkg.config_default(), [0])
    assert ei.value.code == pkg.abi.MPC_ERR_NO_DEVICE and "no CPU fallback" in str(ei.value)


@pytest.mark.gpu
def test_fleet_results_equal_one_engine(pkg):
    import torch
    cfg = pkg.config_default()
    solo = pkg.MpcEngine(cfg, 0)
    fl = pkg.MpcFleet(cfg, [0, 0, 0])                      # three shards on the one GPU
    assert len(fl) == 3
    for n in (0, 1, 2, 1000, 4096):
        st = pkg.generate_states(1002, 17, n)
        want = solo.compute_grf_batch(st).copy() if n else np.zeros(0, dtype=pkg.abi.RESULT_DTYPE)
        got = fl.compute_grf_batch(st)                      # pageable buffers: staged through pinned memory
        assert got.tobytes() == want.tobytes(), n
    # page-locked caller buffers are used in place
    n = 3000
    st = pkg.generate_states(1003, 0, n)
    pin = torch.empty(n * pkg.abi.STATE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    pin_np = pin.numpy().view(pkg.abi.STATE_DTYPE)
    pin_np[:] = st
    out = torch.empty(n * pkg.abi.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(pkg.abi.RESULT_DTYPE)
    fl.compute_grf_batch(pin_np, out)
    assert out.tobytes() == solo.compute_grf_batch(st).tobytes()
    assert fl.kernel_launches() > 0
    # warm-started ticks: every robot keeps its shard, so the fleet follows a single engine's stream
    solo.stream_reset()
    fl.stream_reset()
    for t in range(4):
        s = pkg.generate_stream_states(1006, 0, 500, 46 + t)
        assert fl.stream_step(s).tobytes() == solo.stream_step(s).tobytes(), t
    fl.close()
    solo.close()


@pytest.mark.gpu
def test_fleet_over_every_gpu_of_the_box(pkg, ob):
    import torch
    ndev = torch.cuda.device_count()
    cfg = pkg.config_default()
    fl = pkg.MpcFleet(cfg, list(range(ndev)))
    n = 512 * ndev + 3
    st = pkg.generate_states(1003, 5, n)
    got = fl.compute_grf_batch(st)
    ref = ob.mpc_compute_grf(cfg, st[::7])
    assert np.array_equal(got["iters"][::7], ref["iters"]) and (got["status"] == 1).all()
    den = np.maximum(np.linalg.norm(ref["grf"], axis=1), 1.0)
    assert (np.linalg.norm(got["grf"][::7].astype(np.float64) - ref["grf"], axis=1) / den).max() <= 1e-3
    fl.close()
    with pytest.raises(pkg.MpcError):
        pkg.MpcFleet(cfg, [ndev])                          # no such device
