"""N>1 host logic on CPU: two gloo ranks shard a batch, each fills its shard's result records,
rank 0 gathers.  No data-path collective exists; this covers the partition and
the final gather the multi-GPU bench uses."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_results(pkg, states, first_index):
    """Stand-in for the device solve (CPU test): a result that is a pure function of the state."""
    out = np.zeros(len(states), dtype=pkg.abi.RESULT_DTYPE)
    out["grf"] = np.concatenate([states["foot_pos_abs"][:, :9], states["euler"]], axis=1)
    out["iters"] = np.arange(first_index, first_index + len(states))
    out["status"] = 1
    return out


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import go1_qp_mpc_controller_b200 as pkg
    from go1_qp_mpc_controller_b200.sharding import gather_results, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    states = pkg.generate_states(1003, lo, hi - lo)      # counter-based: no broadcast needed
    local = _fake_results(pkg, states, lo)
    full = gather_results(local, n)
    if rank == 0:
        q.put(full.tobytes())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 64])
def test_two_rank_shard_and_gather(pkg, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=120), dtype=pkg.abi.RESULT_DTYPE)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_results(pkg, pkg.generate_states(1003, 0, n), 0)
    assert got.tobytes() == ref.tobytes()


def test_shard_ranges_partition(pkg):
    from go1_qp_mpc_controller_b200.sharding import owner_of, shard_range
    for n in (0, 1, 5, 37, 4096, 65536, 1_000_000):
        for world in (1, 2, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    assert owner_of(0, 10, 3) == 0 and owner_of(9, 10, 3) == 2 and owner_of(4, 10, 3) == 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)
