"""Host-side logic of the sharded (multi-GPU) pool on CPU: the slice bookkeeping and the
torch.distributed plumbing, exercised with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_plan():
    import stochy_jl_b200 as sb
    sd = sb.dist
    p = sd.shard_plan(1, 4, 1 << 14)
    assert p.n_global == 1 << 16 and p.first_particle == 1 << 14 and p.first_tile == 8 and p.tiles_local == 8
    # slices tile the global index space without gaps
    nxt = 0
    for r in range(4):
        q = sd.shard_plan(r, 4, 1 << 14)
        assert q.first_particle == nxt
        nxt += q.n_local
    assert nxt == p.n_global
    assert sd.owner_of(3 * (1 << 14) + 5, 1 << 14) == (3, 5)
    for bad in [(0, 3, 1 << 14), (2, 2, 1 << 14), (0, 2, 3000), (0, 2, 1024), (0, 8, 1 << 25)]:
        with pytest.raises(ValueError):
            sd.shard_plan(*bad)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import stochy_jl_b200 as sb
        sd = sb.dist
        assert sd.rank_world() == (rank, world)
        mine = bytes([(rank * 37 + i) % 256 for i in range(64)])
        allh = sd.exchange_handles(mine)
        ok = len(allh) == world and all(allh[r] == bytes([(r * 37 + i) % 256 for i in range(64)]) for r in range(world))
        sd.barrier()
        mx = sd.max_over_ranks(10.0 + rank)
        # the per-rank shares of the sharded PMCMC counts (each rank counts the particles it owns) summed over the group
        import numpy as np
        share = np.arange(12, dtype=np.int64).reshape(3, 4) * (rank + 1)
        tot = sd.sum_over_ranks(share)
        ok = ok and tot.dtype == np.int64 and np.array_equal(tot, np.arange(12).reshape(3, 4) * sum(range(1, world + 1)))
        q.put((rank, ok, mx))
    finally:
        dist.destroy_process_group()


def test_handle_exchange_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, mx in res:
        assert ok, "rank %d saw wrong handles" % rank
        assert mx == 11.0          # max over ranks
