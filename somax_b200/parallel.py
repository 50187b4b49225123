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


# ----------------------------------------------------------------------------------------
# One grid over several GPUs: y-slab decomposition of the QG model (BASELINE config 4)
# ----------------------------------------------------------------------------------------
def slab_rows(ny: int, rank: int, world: int) -> range:
    """Interior rows (0-based, of ny) owned by `rank`: equal contiguous blocks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    if ny % world:
        raise ValueError(f"ny ({ny}) must be divisible by the number of slabs ({world})")
    n = ny // world
    return range(rank * n, (rank + 1) * n)


def slab_window(ny: int, rank: int, world: int) -> slice:
    """Rows of the global (Ny = ny + 2) array that form rank's window: its rows plus one row
    either side (the physical ring on the edge ranks, the neighbour's row elsewhere)."""
    r = slab_rows(ny, rank, world)
    return slice(r.start, r.stop + 2)


def slab_owned(ny: int, rank: int, world: int) -> slice:
    """Rows of the global array a rank is authoritative for when slabs are merged: its interior
    rows, plus the ring row on the edge ranks."""
    r = slab_rows(ny, rank, world)
    lo = r.start + 1 - (1 if rank == 0 else 0)
    hi = r.stop + 1 + (1 if rank == world - 1 else 0)
    return slice(lo, hi)


def split_slabs(q, world: int):
    """Global (nl, Ny, Nx) array -> list of per-rank windows (views, halo rows included)."""
    ny = q.shape[-2] - 2
    return [q[..., slab_window(ny, r, world), :] for r in range(world)]


def merge_slabs(slabs, out=None):
    """Inverse of split_slabs: every rank contributes the rows it owns."""
    world = len(slabs)
    nyl = slabs[0].shape[-2] - 2
    ny = nyl * world
    if out is None:
        shape = tuple(slabs[0].shape[:-2]) + (ny + 2, slabs[0].shape[-1])
        if hasattr(slabs[0], "new_empty"):
            out = slabs[0].new_empty(shape)
        else:
            import numpy as np
            out = np.empty(shape, dtype=slabs[0].dtype)
    for r, s in enumerate(slabs):
        own = slab_owned(ny, r, world)
        w0 = slab_window(ny, r, world).start
        out[..., own, :] = s[..., own.start - w0:own.stop - w0, :]
    return out


def gather_blobs(blob: bytes, world: int) -> bytes:
    """All-gather one opaque byte blob per rank (the CUDA-IPC handles of a slab); returns the
    blobs concatenated in rank order.  NCCL groups stage through device memory, gloo on the host."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return bytes(blob)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != world:
        raise RuntimeError(f"torch.distributed must be initialised with world size {world}")
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    parts = [torch.empty(len(blob), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


class SlabQG:
    """A `BaroclinicQG` / `BarotropicQG` model stepped on y-slabs over several GPUs
    (`libsomax_b200`'s `somax_b200_qgs_*`; reference path: core/model.py:53-88 over
    qg/baroclinic.py:135-195).

    * ``SlabQG(model, world)`` inside a torchrun job (``torch.distributed`` initialised, one
      process per GPU): this process holds slab ``rank``; the CUDA-IPC blobs of the peers are
      all-gathered once at construction.  ``integrate_slab`` advances this rank's window.
    * ``SlabQG(model, world, local=True)``: every slab lives in this process on the current GPU
      (validation of the decomposition); ``integrate`` takes and returns the global array.
    """

    def __init__(self, model, world: int, local: bool = False, rank: int | None = None):
        import ctypes as C

        import numpy as np

        from . import _lib
        eng = model._engine
        self.model, self.world, self.local = model, int(world), bool(local)
        self.dtype = np.dtype(eng.dtype)
        self.nl, self.ny, self.nx = eng.nl, eng.ny, eng.nx
        if local:
            self.rank, first, nlocal = 0, 0, self.world
        else:
            import torch.distributed as dist
            if rank is None:
                rank = dist.get_rank() if dist.is_initialized() else 0
            self.rank, first, nlocal = int(rank), int(rank), 1
        h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.somax_b200_qgs_create(
            C.byref(h), _lib.F32 if self.dtype == np.float32 else _lib.F64, eng.nl, eng.ny, eng.nx,
            eng.dx, eng.dy, eng.Cl2m.ctypes.data, eng.Cm2l.ctypes.data, eng.lambdas.ctypes.data,
            eng.beta_y.ctypes.data, eng.wind.ctypes.data, self.world, first, nlocal, eng.spec))
        self._h = h
        if not local:
            self._attach()

    def _attach(self):
        import ctypes as C

        from . import _lib
        L = _lib.lib()
        nb = int(L.somax_b200_qgs_export_bytes())
        mine = (C.c_ubyte * nb)()
        _lib.check(L.somax_b200_qgs_export(self._h, mine))
        raw = gather_blobs(bytes(mine), self.world)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        _lib.check(L.somax_b200_qgs_attach(self._h, buf))

    def window(self):
        return slab_window(self.ny, self.rank, self.world)

    def _steps(self, slabs, n_steps, dt, dt_last):
        import ctypes as C

        from . import _lib
        from .core import stream_ptr
        from .models.qg import _params_struct
        p = _params_struct(self.model.params, self.model._H0)
        ptrs = (C.c_void_p * len(slabs))(*[s.data_ptr() for s in slabs])
        _lib.check(_lib.lib().somax_b200_qgs_steps(self._h, ptrs, int(n_steps), float(dt),
                                                   float(dt_last), C.byref(p), stream_ptr()))

    def check_peers(self):
        """Synchronise and raise if a barrier gave up waiting for a peer."""
        import ctypes as C

        import torch

        from . import _lib
        torch.cuda.current_stream().synchronize()
        n = C.c_int(0)
        _lib.check(_lib.lib().somax_b200_qgs_status(self._h, C.byref(n)))
        if n.value:
            raise _lib.SomaxB200Error(f"slab group: {n.value} barrier(s) timed out waiting for a peer")

    def integrate(self, q0, t0, t1, dt):
        """local=True: Tsit5-advance the global (nl, Ny, Nx) array from t0 to t1; returns the
        same flavour (numpy in -> numpy out)."""
        import numpy as np
        import torch

        from .core import step_plan, torch_dtype
        if not self.local:
            raise RuntimeError("integrate() takes the global array: use local=True, or integrate_slab()")
        n, rem = step_plan(t0, t1, dt)
        was_numpy = not isinstance(q0, torch.Tensor)
        q = torch.as_tensor(np.asarray(q0) if was_numpy else q0).to("cuda", dtype=torch_dtype(self.dtype))
        if q.dim() == 2:
            q = q[None]
        slabs = [s.contiguous().clone() for s in split_slabs(q, self.world)]
        self._steps(slabs, n, dt, rem)
        self.check_peers()
        out = merge_slabs(slabs)
        if np.ndim(q0) == 2:
            out = out[0]
        return out.cpu().numpy() if was_numpy else out

    def integrate_slab(self, slab, t0, t1, dt, check=True):
        """One process per GPU: advance this rank's window (CUDA tensor (nl, ny/world + 2, Nx),
        halo rows valid) in place.  Collective over the slab group."""
        from .core import step_plan
        n, rem = step_plan(t0, t1, dt)
        self.advance_slab(slab, n, dt, rem, check=check)
        return slab

    def advance_slab(self, slab, n_steps, dt, dt_last=0.0, check=True):
        """``check`` (default): synchronise the stream afterwards and raise if a flag barrier gave up
        waiting for a peer (the watchdog sets an error word instead of hanging the GPU; the data
        of such a call is invalid).  ``check=False`` leaves that to a later ``check_peers()``."""
        if self.local:
            raise RuntimeError("advance_slab() is for one-slab-per-process groups")
        if not (slab.is_cuda and slab.is_contiguous()):
            raise ValueError("slab must be a contiguous CUDA tensor")
        self._steps([slab], n_steps, dt, dt_last)
        if check:
            self.check_peers()

    def close(self):
        from . import _lib
        if getattr(self, "_h", None) is not None:
            _lib.lib().somax_b200_qgs_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class SlabSWM:
    """A `MultilayerShallowWater2D` / `NonlinearShallowWater2D` model stepped on y-slabs over several
    GPUs (`libsomax_b200`'s `somax_b200_swms_*`; reference path: core/model.py:53-88 over
    swm/multilayer.py:150-223).  Halo exchange only: one row of (h, u, v) per neighbour and
    right-hand-side evaluation, pushed into the neighbour's memory (CUDA IPC over NVLink).

    Same two uses as `SlabQG`: one slab per process inside a torchrun job (`advance_slab` /
    `integrate_slab` on this rank's windows), or `local=True` with every slab in this process on the
    current GPU (validation; `integrate` takes and returns the global arrays).
    """

    def __init__(self, model, world: int, local: bool = False, rank: int | None = None):
        import ctypes as C

        import numpy as np

        from . import _lib
        self.model, self.world, self.local = model, int(world), bool(local)
        self.dtype = np.dtype(model.dtype)
        g = model.grid
        self.nl, self.ny, self.nx = model._nl, g.Ny - 2, g.Nx - 2
        if model._base_ndim != 3 and model._nl != 1:
            raise ValueError("unexpected model layout")
        if local:
            self.rank, first, nlocal = 0, 0, self.world
        else:
            import torch.distributed as dist
            if rank is None:
                rank = dist.get_rank() if dist.is_initialized() else 0
            self.rank, first, nlocal = int(rank), int(rank), 1
        h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.somax_b200_swms_create(
            C.byref(h), _lib.F32 if self.dtype == np.float32 else _lib.F64, self.nl, self.ny, self.nx, g.dx, g.dy,
            _lib.BC_PERIODIC if model._bc == "periodic" else _lib.BC_WALL, model._g.ctypes.data,
            model._f.ctypes.data, model._wx.ctypes.data, model._wy.ctypes.data, self.world, first, nlocal,
            model._spec))
        self._h = h
        if not local:
            nb = int(L.somax_b200_swms_export_bytes())
            mine = (C.c_ubyte * nb)()
            _lib.check(L.somax_b200_swms_export(self._h, mine))
            raw = gather_blobs(bytes(mine), self.world)
            buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
            _lib.check(L.somax_b200_swms_attach(self._h, buf))

    def window(self):
        return slab_window(self.ny, self.rank, self.world)

    def _steps(self, hs, us, vs, n_steps, dt, dt_last):
        import ctypes as C

        from . import _lib
        from .core import stream_ptr
        p = self.model._pstruct()
        arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        _lib.check(_lib.lib().somax_b200_swms_steps(self._h, arr(hs), arr(us), arr(vs), int(n_steps), float(dt),
                                                    float(dt_last), C.byref(p), stream_ptr()))

    def check_peers(self):
        """Synchronise and raise if a barrier gave up waiting for a peer."""
        import ctypes as C

        import torch

        from . import _lib
        torch.cuda.current_stream().synchronize()
        n = C.c_int(0)
        _lib.check(_lib.lib().somax_b200_swms_status(self._h, C.byref(n)))
        if n.value:
            raise _lib.SomaxB200Error(f"slab group: {n.value} barrier(s) timed out waiting for a peer")

    def integrate(self, state0, t0, t1, dt):
        """local=True: Tsit5-advance the global (h, u, v) arrays from t0 to t1; returns a state of
        the same class (numpy in -> numpy out)."""
        import numpy as np
        import torch

        from .core import step_plan, torch_dtype
        if not self.local:
            raise RuntimeError("integrate() takes the global arrays: use local=True, or integrate_slab()")
        n, rem = step_plan(t0, t1, dt)
        was_numpy = not isinstance(state0.h, torch.Tensor)
        single = np.ndim(state0.h) == 2
        parts = []
        for f in ("h", "u", "v"):
            a = getattr(state0, f)
            t = torch.as_tensor(np.asarray(a) if was_numpy else a).to("cuda", dtype=torch_dtype(self.dtype))
            if single:
                t = t[None]
            parts.append([s.contiguous().clone() for s in split_slabs(t, self.world)])
        self._steps(parts[0], parts[1], parts[2], n, dt, rem)
        self.check_peers()
        outs = []
        for p in parts:
            o = merge_slabs(p)
            if single:
                o = o[0]
            outs.append(o.cpu().numpy() if was_numpy else o)
        return type(state0)(h=outs[0], u=outs[1], v=outs[2])

    def advance_slab(self, h, u, v, n_steps, dt, dt_last=0.0, check=True):
        """One process per GPU: advance this rank's windows (contiguous CUDA tensors
        (nl, ny/world + 2, Nx), halo rows valid) in place.  Collective over the slab group."""
        if self.local:
            raise RuntimeError("advance_slab() is for one-slab-per-process groups")
        for t in (h, u, v):
            if not (t.is_cuda and t.is_contiguous()):
                raise ValueError("slabs must be contiguous CUDA tensors")
        self._steps([h], [u], [v], n_steps, dt, dt_last)
        if check:
            self.check_peers()

    def integrate_slab(self, h, u, v, t0, t1, dt, check=True):
        from .core import step_plan
        n, rem = step_plan(t0, t1, dt)
        self.advance_slab(h, u, v, n, dt, rem, check=check)
        return h, u, v

    def close(self):
        from . import _lib
        if getattr(self, "_h", None) is not None:
            _lib.lib().somax_b200_swms_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
