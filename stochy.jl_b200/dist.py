"""torch.distributed plumbing for sharded pools (one process per GPU).

Only host-side bookkeeping lives here: which rank owns which slice of the global particle index
space, and the exchange of the 64-byte CUDA-IPC handles of the per-GPU arenas.  The data path itself
uses no collective library: after the exchange the CUDA kernels read peer memory over NVLink
(see DESIGN.md "Multi-GPU").  Works with the nccl backend (GPU tensors) and with gloo (CPU tensors),
which is what the CPU tests use.
"""
from collections import namedtuple

TILE = 2048
MAX_WORLD = 8
MAX_GLOBAL = 1 << 27

Shard = namedtuple("Shard", "rank world n_local n_global first_particle first_tile tiles_local")


def shard_plan(rank, world, n_local):
    """Slice of the global pool owned by `rank`: particles [rank * n_local, (rank + 1) * n_local)."""
    if world < 1 or world > MAX_WORLD or world & (world - 1):
        raise ValueError("world must be 1, 2, 4 or 8")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    if n_local < TILE or n_local & (n_local - 1):
        raise ValueError("n_local must be a power of two >= %d" % TILE)
    if n_local * world > MAX_GLOBAL:
        raise ValueError("world x n_local must not exceed 2^27 particles")
    return Shard(rank, world, n_local, n_local * world, rank * n_local, rank * (n_local // TILE), n_local // TILE)


def owner_of(particle, n_local):
    """(rank, local index) of a global particle index -- the address computation the kernels do."""
    return particle // n_local, particle % n_local


def rank_world(group=None):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _device_for(group):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def exchange_handles(mine, group=None):
    """all_gather of one fixed-size byte string per rank; returns the list in rank order."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world(group)
    if world == 1:
        return [bytes(mine)]
    dev = _device_for(group)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().tolist()) for o in out]


def barrier(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


def max_over_ranks(value, group=None):
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world(group)
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(arr, group=None):
    """Integer histograms kept per rank (each rank counts the particles it owns) summed over the group."""
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world = rank_world(group)
    if world == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device_for(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
