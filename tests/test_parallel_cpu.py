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


def test_slab_geometry_helpers():
    """Windows / owned rows of the y-slab decomposition (host logic of parallel.SlabQG)."""
    import numpy as np
    from somax_b200.parallel import merge_slabs, slab_owned, slab_rows, slab_window, split_slabs
    ny = 24
    for world in (1, 2, 3, 4, 8):
        rows = [slab_rows(ny, r, world) for r in range(world)]
        assert [j for r in rows for j in r] == list(range(ny))
        owned = [slab_owned(ny, r, world) for r in range(world)]
        assert owned[0].start == 0 and owned[-1].stop == ny + 2           # ring rows belong to the edge ranks
        assert all(a.stop == b.start for a, b in zip(owned, owned[1:]))    # a partition of the Ny rows
        for r in range(world):
            w = slab_window(ny, r, world)
            assert w.stop - w.start == ny // world + 2
            assert w.start <= owned[r].start and owned[r].stop <= w.stop
        q = np.random.default_rng(world).standard_normal((3, ny + 2, 7))
        slabs = [s.copy() for s in split_slabs(q, world)]
        for r in range(world - 1):                                         # neighbours share two rows
            assert np.array_equal(slabs[r][:, -2:], slabs[r + 1][:, :2])
        # halo rows are not authoritative: garbage there must not reach the merged array
        for r, s in enumerate(slabs):
            if r > 0:
                s[:, 0] = np.nan
            if r < world - 1:
                s[:, -1] = np.nan
        assert np.array_equal(merge_slabs(slabs), q)
    with pytest.raises(ValueError):
        slab_rows(25, 0, 2)
    with pytest.raises(ValueError):
        slab_rows(24, 2, 2)


def _blob_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the set-up exchange of SlabQG._attach: every rank contributes an opaque byte blob, all ranks
    # end up with the blobs in rank order
    from somax_b200.parallel import gather_blobs
    q.put((rank, gather_blobs(bytes([rank + 1] * 216), world)))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_blob_all_gather_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_blob_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes([1] * 216 + [2] * 216)
    assert all(blob == want for _, blob in res)
