"""Multi-GPU plumbing: the batch shards trivially (every QP is independent).

One process per GPU (torch.distributed, NCCL on GPUs / gloo in CPU tests).  State i of a
batch of n goes to rank floor(i * world / n) as contiguous index ranges;
there is NO collective on the data path.  The only exchange is the final gather of the
64 B result records to rank 0.
"""
import numpy as np

from . import abi


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of the n problems owned by `rank`; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request")
    lo = (n * rank + world - 1) // world if False else -(-n * rank // world)
    hi = -(-n * (rank + 1) // world)
    return lo, hi


def owner_of(i, n, world):
    """Rank that owns problem i."""
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        if lo <= i < hi:
            return r
    raise ValueError("index out of range")


def gather_results(local, n_total, group=None, dst=0):
    """Final gather of MpcResult records to rank `dst` (returns None elsewhere).

    `local` is this rank's RESULT_DTYPE array.  Uses all_gather on byte tensors padded to
    the largest shard, on the device of the process group's backend (NCCL -> CUDA tensors
    over NVLink, gloo -> CPU tensors).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert local.dtype == abi.RESULT_DTYPE
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_len = max(hi - lo for lo, hi in sizes)
    pad = np.zeros(max_len, dtype=abi.RESULT_DTYPE)
    pad[:len(local)] = local
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(pad.view(np.uint8).copy()).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    if rank != dst:
        return None
    full = np.zeros(n_total, dtype=abi.RESULT_DTYPE)
    for r, (lo, hi) in enumerate(sizes):
        full[lo:hi] = outs[r].cpu().numpy().view(abi.RESULT_DTYPE)[:hi - lo]
    return full
