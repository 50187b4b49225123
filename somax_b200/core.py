"""Model contract, core types, stratification and modal transform.

Host-side mirror of the reference's ``somax/_src/core`` for the hot path
(core/model.py:12-95, core/types.py:8-36, core/transforms.py:13-224).  Arrays are
``torch`` CUDA tensors (device memory plumbing); numpy inputs are staged through pinned
host memory and results come back as numpy.  Nothing here computes the model physics: the
right-hand sides, the PV inversion and the Tsit5 loop run in ``libsomax_b200.so``.
"""
from __future__ import annotations

import abc
import math
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Sequence

import numpy as np

from . import _lib

try:  # torch is plumbing only (device buffers, streams)
    import torch
except Exception as _e:  # pragma: no cover
    torch = None
    _torch_error = _e


# ----------------------------------------------------------------------------------------
# core types (core/types.py:8-36)
# ----------------------------------------------------------------------------------------
class State:
    """Base class for model states (dataclass subclasses hold the prognostic arrays)."""


class Params:
    """Differentiable parameters (floats or 0-d arrays)."""


class PhysConsts:
    """Frozen physical constants."""


class Diagnostics:
    """Diagnostic quantities."""


@dataclass(frozen=True)
class Grid:
    """Stand-in for ``finitevolx.ArakawaCGrid2D`` (one ghost ring)."""

    Nx: int
    Ny: int
    Lx: float
    Ly: float
    dx: float
    dy: float

    @staticmethod
    def from_interior(nx: int, ny: int, Lx: float, Ly: float) -> "Grid":
        return Grid(Nx=nx + 2, Ny=ny + 2, Lx=float(Lx), Ly=float(Ly), dx=Lx / nx, dy=Ly / ny)


# ----------------------------------------------------------------------------------------
# stratification + modal transform (core/transforms.py:13-224); setup-time host maths
# ----------------------------------------------------------------------------------------
@dataclass
class StratificationProfile:
    H: np.ndarray
    g_prime: np.ndarray
    rho: np.ndarray | None = None

    @property
    def nl(self) -> int:
        return int(np.shape(self.H)[0])

    @property
    def total_depth(self):
        return float(np.sum(self.H))

    @staticmethod
    def from_layers(H, g_prime, rho=None) -> "StratificationProfile":
        if len(H) != len(g_prime):
            raise ValueError(f"H ({len(H)}) and g_prime ({len(g_prime)}) must have the same length")
        if rho is not None and len(rho) != len(H):
            raise ValueError(f"rho ({len(rho)}) must have the same length as H ({len(H)})")
        return StratificationProfile(H=np.asarray(H, np.float64), g_prime=np.asarray(g_prime, np.float64),
                                     rho=None if rho is None else np.asarray(rho, np.float64))

    @staticmethod
    def from_N2_constant(N2, depth, n_layers, g=9.81, rho0=1025.0) -> "StratificationProfile":
        Hv = depth / n_layers
        H = np.full(n_layers, Hv)
        g_prime = np.concatenate([[g], np.full(n_layers - 1, N2 * Hv)])
        rho = rho0 + (rho0 * N2 * Hv / g) * np.arange(n_layers)
        return StratificationProfile(H=H, g_prime=g_prime, rho=rho)

    @staticmethod
    def from_N2_exponential(N2_surface, scale_depth, depth, n_layers, g=9.81,
                            rho0=1025.0) -> "StratificationProfile":
        Hv = depth / n_layers
        H = np.full(n_layers, Hv)
        zi = -np.arange(1, n_layers) * Hv
        g_prime = np.concatenate([[g], N2_surface * np.exp(zi / scale_depth) * Hv])
        zc = -(np.arange(n_layers) + 0.5) * Hv
        rho = rho0 + np.cumsum(rho0 * N2_surface * np.exp(zc / scale_depth) * Hv / g)
        return StratificationProfile(H=H, g_prime=g_prime, rho=rho)


def build_coupling_matrix(H, g_prime) -> np.ndarray:
    """``finitevolx.build_coupling_matrix`` (MQGeometry convention)."""
    H = np.asarray(H, np.float64)
    g = np.asarray(g_prime, np.float64)
    nl = H.shape[0]
    if nl == 1:
        return np.array([[1.0 / (H[0] * g[0])]])
    A = np.zeros((nl, nl))
    A[0, 0] = 1.0 / (H[0] * g[0]) + 1.0 / (H[0] * g[1])
    A[0, 1] = -1.0 / (H[0] * g[1])
    for k in range(1, nl - 1):
        A[k, k - 1] = -1.0 / (H[k] * g[k])
        A[k, k] = (1.0 / g[k] + 1.0 / g[k + 1]) / H[k]
        A[k, k + 1] = -1.0 / (H[k] * g[k + 1])
    A[-1, -2] = -1.0 / (H[-1] * g[-1])
    A[-1, -1] = 1.0 / (H[-1] * g[-1])
    return A


def decompose_vertical_modes(A, f0):
    """``finitevolx.decompose_vertical_modes``: (rossby_radii, Cl2m, Cm2l)."""
    A = np.asarray(A, np.float64)
    wr, R = np.linalg.eig(A)
    wl, Lm = np.linalg.eig(A.T)
    R = R.real[:, np.argsort(wr.real)]
    Lm = Lm.real[:, np.argsort(wl.real)]
    w = np.sort(wr.real)
    Cl2m = np.diag(1.0 / np.diag(Lm.T @ R)) @ Lm.T
    with np.errstate(divide="ignore", invalid="ignore"):
        radii = 1.0 / (abs(f0) * np.sqrt(np.abs(w)))
    return radii, Cl2m, R


@dataclass
class ModalTransform:
    """core/transforms.py:151-224.  ``eigenvalues`` reproduces ``jnp.linalg.eigh(A)`` (which
    symmetrises its input) exactly as the reference calls it at :194 - see SURVEY 0-8(i)."""

    Cl2m: np.ndarray
    Cm2l: np.ndarray
    eigenvalues: np.ndarray
    rossby_radii: np.ndarray

    @staticmethod
    def from_physics(H, g_prime, f0) -> "ModalTransform":
        A = build_coupling_matrix(H, g_prime)
        radii, Cl2m, Cm2l = decompose_vertical_modes(A, f0)
        ev = np.linalg.eigvalsh(0.5 * (A + A.T))
        return ModalTransform(Cl2m=Cl2m, Cm2l=Cm2l, eigenvalues=ev, rossby_radii=radii)

    @staticmethod
    def from_stratification(strat: StratificationProfile, f0) -> "ModalTransform":
        return ModalTransform.from_physics(strat.H, strat.g_prime, f0)

    def to_modal(self, x):
        return _einsum_lm(self.Cl2m, x)

    def to_layer(self, x):
        return _einsum_lm(self.Cm2l, x)


def _einsum_lm(M, x):
    if torch is not None and isinstance(x, torch.Tensor):
        Mt = torch.as_tensor(M, dtype=x.dtype, device=x.device)
        return torch.einsum("lm,m...->l...", Mt, x)
    x = np.asarray(x)
    return np.einsum("lm,m...->l...", np.asarray(M, x.dtype if x.dtype.kind == "f" else np.float64), x)


# ----------------------------------------------------------------------------------------
# device plumbing
# ----------------------------------------------------------------------------------------
def _require_torch():
    if torch is None:  # pragma: no cover
        raise _lib.SomaxB200Error(f"torch is required for device buffers: {_torch_error}")
    if not torch.cuda.is_available():
        raise _lib.SomaxB200Error("no CUDA device visible: somax_b200 has no CPU fallback")


def torch_dtype(dtype):
    return torch.float32 if np.dtype(dtype) == np.float32 else torch.float64


class DeviceIO:
    """Moves caller arrays to device tensors of the model dtype and results back in the
    caller's flavour:

    * numpy in -> numpy out, staged through pinned buffers that are cached on the model (the
      ``cudaHostAlloc`` of a state-sized buffer costs far more than the copy it serves);
    * pinned CPU ``torch.Tensor`` in -> pinned CPU ``torch.Tensor`` out, no staging copy at all:
      the H2D DMA reads the caller's buffer, the D2H DMA lands in a cached pinned buffer that is
      returned as is (valid until the next call on the same model; ``.clone()`` to keep it);
    * CUDA ``torch.Tensor`` in -> CUDA ``torch.Tensor`` out, nothing leaves the device.
    """

    def __init__(self, dtype, cache=None):
        _require_torch()
        self.dtype = np.dtype(dtype)
        self.tdtype = torch_dtype(dtype)
        self.numpy_out = False
        self.host_out = False
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.cache = cache if cache is not None else {}
        self._nin = 0
        self._nout = 0

    def _pinned(self, role, shape):
        key = (role, tuple(shape), self.tdtype)
        buf = self.cache.get(key)
        if buf is None:
            buf = torch.empty(tuple(shape), dtype=self.tdtype, pin_memory=True)
            self.cache[key] = buf
        return buf

    def to_device(self, x):
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                t = x.to(dtype=self.tdtype).contiguous()
                return t.clone() if t.data_ptr() == x.data_ptr() else t
            self.host_out = True
            if x.is_pinned() and x.dtype == self.tdtype and x.is_contiguous():
                self.h2d_bytes += x.numel() * x.element_size()
                return x.to("cuda", non_blocking=True)
            a = x.to(dtype=self.tdtype)
            pinned = self._pinned(("in", self._nin), a.shape)
            self._nin += 1
            pinned.copy_(a)
            self.h2d_bytes += pinned.numel() * pinned.element_size()
            return pinned.to("cuda", non_blocking=True)
        self.numpy_out = True
        a = np.asarray(x)
        pinned = self._pinned(("in", self._nin), a.shape)
        self._nin += 1
        np.copyto(pinned.numpy(), a, casting="unsafe")
        self.h2d_bytes += pinned.numel() * pinned.element_size()
        return pinned.to("cuda", non_blocking=True)

    def from_device(self, t):
        if not (self.numpy_out or self.host_out):
            return t
        pinned = self._pinned(("out", self._nout), t.shape)
        self._nout += 1
        pinned.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes += pinned.numel() * pinned.element_size()
        return pinned.numpy().copy() if self.numpy_out else pinned


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def scalar(x) -> float:
    if torch is not None and isinstance(x, torch.Tensor):
        return float(x.item())
    return float(np.asarray(x))


# ----------------------------------------------------------------------------------------
# diffrax stand-ins for the pieces ``SomaxModel.integrate`` exposes (core/model.py:53-88)
# ----------------------------------------------------------------------------------------
@dataclass
class SaveAt:
    """``diffrax.SaveAt`` subset: ``t1=True`` (default) or explicit ``ts``."""

    t1: bool = False
    ts: Sequence[float] | None = None
    t0: bool = False


class Tsit5:
    """Marker for the only solver the CUDA path implements (diffrax.Tsit5)."""


class ConstantStepSize:
    """Marker for the only step-size controller the CUDA path implements."""


@dataclass
class ODETerm:
    """``diffrax.ODETerm`` stand-in: ``vf(t, y, args)`` = ``vector_field(BC(y))``."""

    vector_field: Any

    def vf(self, t, y, args=None):
        return self.vector_field(t, y, args)


@dataclass
class Solution:
    """``diffrax.Solution`` subset: ``ts`` (n_save,), ``ys`` state with a leading time axis."""

    ts: np.ndarray
    ys: Any
    stats: dict = field(default_factory=dict)


def step_plan(t0: float, t1: float, dt: float):
    """(n_full, dt_last) of ConstantStepSize with the last step clipped to t1."""
    span = float(t1) - float(t0)
    if span < 0 or not dt > 0:
        raise ValueError("need t1 >= t0 and dt > 0")
    n = int(math.floor(span / dt * (1.0 + 1e-12) + 1e-9))
    rem = span - n * dt
    if rem <= 1e-9 * abs(dt):
        rem = 0.0
    return n, rem


def _own_host(state):
    """Copy of ``state`` whose host torch tensors are owned: the pinned staging buffer a model
    returns is reused by its next call, so several save times would otherwise alias one buffer."""
    kw = {}
    for f in fields(state):
        v = getattr(state, f.name)
        kw[f.name] = v.clone() if (torch is not None and isinstance(v, torch.Tensor) and not v.is_cuda) else v
    return type(state)(**kw)


def _stack_states(cls, states):
    names = [f.name for f in fields(cls)]
    vals = {}
    for nm in names:
        parts = [getattr(s, nm) for s in states]
        if len(parts) == 1:
            vals[nm] = parts[0][None]          # a view: no state-sized copy for the usual t1 save
        elif isinstance(parts[0], np.ndarray):
            vals[nm] = np.stack(parts)
        else:
            vals[nm] = torch.stack(parts)
    return cls(**vals)


class SomaxModel(abc.ABC):
    """The somax model contract (core/model.py:12-95)."""

    @abc.abstractmethod
    def vector_field(self, t, state, args=None):
        ...

    @abc.abstractmethod
    def apply_boundary_conditions(self, state):
        ...

    @abc.abstractmethod
    def _advance(self, state, n_steps: int, dt: float, dt_last: float, resume: bool = False):
        """Tsit5-advance ``state`` on the device.  ``resume=False``: boundary conditions are applied
        to it first (the start of an integration); ``resume=True``: continue from a state an
        earlier call returned, un-projected."""

    def build_terms(self) -> ODETerm:
        def _rhs(t, state, args=None):
            state = self.apply_boundary_conditions(state)
            return self.vector_field(t, state, args)

        return ODETerm(_rhs)

    def integrate(self, state0, t0: float, t1: float, dt: float, **kw) -> Solution:
        """Forward integration (core/model.py:53-88).  Supports ``saveat`` (``t0`` / ``ts`` / ``t1``),
        ``max_steps``; ``solver`` must be Tsit5 and ``stepsize_controller`` constant.

        As in the reference, the boundary conditions are applied to ``state0`` once
        (core/model.py:62) and afterwards only inside the right-hand side: the states saved at
        intermediate times carry the drifting ghost ring, and stepping continues from them
        un-projected.  diffrax reaches save times that are not on the step grid ``t0 + k dt`` by
        dense-output interpolation, which the CUDA path does not implement: such ``ts`` raise."""
        solver = kw.pop("solver", None)
        if solver is not None and type(solver).__name__ != "Tsit5":
            raise NotImplementedError("the CUDA path implements diffrax.Tsit5 only")
        ctrl = kw.pop("stepsize_controller", None)
        if ctrl is not None and type(ctrl).__name__ != "ConstantStepSize":
            raise NotImplementedError("the CUDA path implements ConstantStepSize only")
        saveat = kw.pop("saveat", None) or SaveAt(t1=True)
        max_steps = kw.pop("max_steps", 4096)
        if kw:
            raise TypeError(f"unsupported integrate() arguments: {sorted(kw)}")
        sub = getattr(saveat, "subs", None)          # a real diffrax.SaveAt keeps its fields in .subs
        src = sub if (sub is not None and not isinstance(sub, (list, tuple))) else saveat
        ts_attr = getattr(src, "ts", None)
        want_t0, want_t1 = bool(getattr(src, "t0", False)), bool(getattr(src, "t1", False))
        ts_list = [float(t) for t in (np.asarray(ts_attr).reshape(-1).tolist() if ts_attr is not None else [])]
        if ts_attr is None and not (want_t0 or want_t1):
            want_t1 = True
        t0, t1, dt = float(t0), float(t1), float(dt)
        n_total, rem_total = step_plan(t0, t1, dt)
        if max_steps is not None and n_total + (1 if rem_total > 0 else 0) > max_steps:
            raise RuntimeError(
                f"max_steps ({max_steps}) reached: {n_total + (rem_total > 0)} steps are needed")
        prev = t0
        for t in ts_list:
            k = (t - t0) / dt
            on_grid = abs(k - round(k)) <= 1e-9 * max(1.0, abs(k)) or abs(t - t1) <= 1e-9 * max(abs(t1), abs(dt))
            if t < prev or t > t1 * (1 + 1e-12) + 1e-300 or not on_grid:
                raise ValueError(
                    f"saveat.ts must be increasing, inside [t0, t1] and on the step grid t0 + k*dt (or t1): {t} is "
                    "not; diffrax interpolates such times from its dense output, the CUDA path does not")
            prev = t
        save_ts = ([t0] if want_t0 else []) + ts_list + ([t1] if want_t1 else [])
        outs = []
        cur, tcur, started = state0, t0, False
        many = len(save_ts) > 1
        for ts_ in save_ts:
            n, rem = step_plan(tcur, ts_, dt)
            if not started or n > 0 or rem > 0:
                # the first call applies the BCs to state0 (also when it takes no step: SaveAt(t0=True))
                cur = self._advance(cur, n, dt, rem, resume=started)
                started = True
            tcur = ts_
            outs.append(_own_host(cur) if many else cur)
        ys = _stack_states(type(outs[0]), outs)
        return Solution(ts=np.asarray(save_ts), ys=ys, stats={"num_steps": n_total + (rem_total > 0)})

    def diagnose(self, state):
        return {}
