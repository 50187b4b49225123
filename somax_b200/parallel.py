"""Multi-GPU plumbing: one process per GPU, ensemble members sharded by rank.

The member axis is embarrassingly parallel: no data-path collective.  The only
collectives are a barrier and the MAX all-reduce that turns per-rank device times into the job
time.  Works with NCCL on GPUs and with gloo on CPUs (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

import os


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_members(n_members: int, rank: int, world: int) -> range:
    """Contiguous block of ensemble members owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or n_members < 0:
        raise ValueError("need 0 <= rank < world and n_members >= 0")
    base, extra = divmod(n_members, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def job_time_ms(local_ms: float, device=None) -> float:
    """MAX over ranks of the per-rank device time (the job finishes with its slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms)
    t = torch.tensor([local_ms], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_member_scalars(local_values, device=None):
    """All-gather per-member diagnostics scalars (e.g. energies) to every rank, in member order."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(local_values, dtype=torch.float64, device=device or "cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in sizes]
    dist.all_gather(outs, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])
