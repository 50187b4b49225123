"""CPU: the N > 1 host logic under torch.distributed gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from somax_b200.parallel import shard_members


def test_shard_members_partition():
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            parts = [shard_members(n, r, w) for r in range(w)]
            flat = [m for p in parts for m in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_members(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from somax_b200.parallel import env_world, gather_member_scalars, job_time_ms
    assert env_world() == (rank, rank, world)
    mine = shard_members(5, rank, world)
    t = job_time_ms(10.0 + 5.0 * rank)                    # slowest rank defines the job time
    g = gather_member_scalars([[float(m), float(m) ** 2] for m in mine])
    dist.barrier()
    q.put((rank, t, g.tolist()))
    dist.destroy_process_group()


def test_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t, g in res:
        assert t == 15.0
        assert g == [[float(m), float(m) ** 2] for m in range(5)]
