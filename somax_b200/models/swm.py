"""Nonlinear shallow-water models (single- and multi-layer) on the B200 path.

Mirror of somax/_src/models/swm/multilayer.py:29-410 and swm/nonlinear_2d.py:25-318: same
class names, fields, methods and ``create`` signatures; the arithmetic runs in
``libsomax_b200.so`` (fused single-pass RHS + Tsit5 epilogue).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib
from ..core import (DeviceIO, Diagnostics, Grid, ModalTransform, Params, PhysConsts, SomaxModel,
                    State, StratificationProfile, scalar, stream_ptr, torch)


@dataclass
class MultilayerSW2DState(State):
    """h, u, v: (nl, Ny, Nx) each (or with a leading members axis)."""
    h: object
    u: object
    v: object


@dataclass
class NonlinearSW2DState(State):
    """h, u, v: (Ny, Nx) each (or with a leading members axis)."""
    h: object
    u: object
    v: object


@dataclass
class MultilayerSW2DParams(Params):
    lateral_viscosity: object
    bottom_drag: object
    wind_amplitude: object


@dataclass
class NonlinearSW2DParams(Params):
    lateral_viscosity: object
    bottom_drag: object
    wind_amplitude: object


@dataclass(frozen=True)
class MultilayerSW2DPhysConsts(PhysConsts):
    gravity: float = 9.81
    f0: float = 1e-4
    beta: float = 0.0
    n_layers: int = 2


@dataclass(frozen=True)
class NonlinearSW2DPhysConsts(PhysConsts):
    gravity: float = 9.81
    f0: float = 1e-4
    beta: float = 0.0
    H0: float = 100.0


@dataclass
class MultilayerSW2DDiagnostics(Diagnostics):
    energy: object
    total_energy: object
    enstrophy: object
    total_enstrophy: object
    potential_vorticity: object
    relative_vorticity: object
    kinetic_energy_field: object
    nonfinite: object = None


@dataclass
class NonlinearSW2DDiagnostics(Diagnostics):
    energy: object
    enstrophy: object
    potential_vorticity: object
    relative_vorticity: object
    kinetic_energy_field: object
    nonfinite: object = None


class _SWMBase(SomaxModel):
    _base_ndim = 3
    _state_cls = MultilayerSW2DState

    def _setup(self, dtype, nl, grid, bc, g_prime, f_field, wind_x, wind_y, H0, spec):
        self.dtype = np.dtype(dtype)
        self._nl, self._bc, self._H0, self._spec = nl, bc, float(H0), spec
        self._g = np.ascontiguousarray(g_prime, np.float64)
        self._f = np.ascontiguousarray(f_field, np.float64)
        self._wx = np.ascontiguousarray(wind_x, np.float64)
        self._wy = np.ascontiguousarray(wind_y, np.float64)
        self._handles = {}

    def _handle(self, batch):
        if batch not in self._handles:
            h = C.c_void_p()
            g = self.grid
            _lib.check(_lib.lib().somax_b200_swm_create(
                C.byref(h), _lib.F32 if self.dtype == np.float32 else _lib.F64, batch, self._nl,
                g.Ny - 2, g.Nx - 2, g.dx, g.dy,
                _lib.BC_PERIODIC if self._bc == "periodic" else _lib.BC_WALL,
                self._g.ctypes.data, self._f.ctypes.data, self._wx.ctypes.data, self._wy.ctypes.data,
                self._spec))
            self._handles[batch] = h
        return self._handles[batch]

    def close(self):
        for h in self._handles.values():
            _lib.lib().somax_b200_swm_destroy(h)
        self._handles = {}

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _batch(self, a):
        shp = tuple(a.shape)
        g = self.grid
        core = (self._nl, g.Ny, g.Nx) if self._base_ndim == 3 else (g.Ny, g.Nx)
        if shp == core:
            return 1
        if len(shp) == self._base_ndim + 1 and shp[1:] == core:
            return shp[0]
        raise ValueError(f"state shape {shp} does not match the model grid {core}")

    def _dev(self, state):
        io = DeviceIO(self.dtype, self.__dict__.setdefault("_io_cache", {}))
        h, u, v = io.to_device(state.h), io.to_device(state.u), io.to_device(state.v)
        return io, h, u, v, self._handle(self._batch(h))

    def _pstruct(self):
        p = self.params
        return _lib.ParamsStruct(scalar(p.lateral_viscosity), scalar(p.bottom_drag),
                                 scalar(p.wind_amplitude), self._H0)

    def vector_field(self, t, state, args=None):
        io, h, u, v, hd = self._dev(state)
        dh, du, dv = torch.empty_like(h), torch.empty_like(u), torch.empty_like(v)
        p = self._pstruct()
        _lib.check(_lib.lib().somax_b200_swm_rhs(hd, h.data_ptr(), u.data_ptr(), v.data_ptr(),
                                                 dh.data_ptr(), du.data_ptr(), dv.data_ptr(),
                                                 C.byref(p), 0, stream_ptr()))
        return self._state_cls(h=io.from_device(dh), u=io.from_device(du), v=io.from_device(dv))

    def apply_boundary_conditions(self, state):
        io, h, u, v, hd = self._dev(state)
        ho, uo, vo = torch.empty_like(h), torch.empty_like(u), torch.empty_like(v)
        _lib.check(_lib.lib().somax_b200_swm_apply_bc(hd, h.data_ptr(), u.data_ptr(), v.data_ptr(),
                                                      ho.data_ptr(), uo.data_ptr(), vo.data_ptr(),
                                                      stream_ptr()))
        return self._state_cls(h=io.from_device(ho), u=io.from_device(uo), v=io.from_device(vo))

    def _advance(self, state, n_steps, dt, dt_last, resume=False):
        io, h, u, v, hd = self._dev(state)
        p = self._pstruct()
        fn = _lib.lib().somax_b200_swm_resume if resume else _lib.lib().somax_b200_swm_steps
        _lib.check(fn(hd, h.data_ptr(), u.data_ptr(), v.data_ptr(), int(n_steps), float(dt), float(dt_last),
                      C.byref(p), stream_ptr()))
        self.last_io = io
        return self._state_cls(h=io.from_device(h), u=io.from_device(u), v=io.from_device(v))

    def diag_scalars(self, state):
        """(energy[nl], potential enstrophy[nl], nonfinite) from the fused device reduction."""
        io, h, u, v, hd = self._dev(state)
        batch, nl = self._batch(h), self._nl
        out = torch.zeros((batch, 3 * nl + 1), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().somax_b200_swm_diag(hd, h.data_ptr(), u.data_ptr(), v.data_ptr(),
                                                  out.data_ptr(), stream_ptr()))
        vals = out.cpu().numpy()
        ke, h2, ens, bad = vals[:, :nl], vals[:, nl:2 * nl], vals[:, 2 * nl:3 * nl], vals[:, 3 * nl]
        energy = ke + 0.5 * self._g[None, :] * h2
        if len(tuple(state.h.shape)) == self._base_ndim:
            energy, ens, bad = energy[0], ens[0], bad[0]
        return energy, ens, bad

    def _diag_fields(self, state):
        """q, zeta, ke fields with interior-only semantics (swm/multilayer.py:231-234);
        elementwise device ops, not on the hot path."""
        h, u, v = state.h, state.u, state.v
        xp = np if isinstance(h, np.ndarray) else torch
        dx, dy = self.grid.dx, self.grid.dy
        I = (Ellipsis, slice(1, -1), slice(1, -1))
        ke = xp.zeros_like(h)
        zeta = xp.zeros_like(h)
        q = xp.zeros_like(h)
        ke[I] = 0.5 * (0.5 * (u[I] ** 2 + u[..., 1:-1, :-2] ** 2) + 0.5 * (v[I] ** 2 + v[..., :-2, 1:-1] ** 2))
        zeta[I] = (v[..., 1:-1, 2:] - v[I]) / dx - (u[..., 2:, 1:-1] - u[I]) / dy
        f = self.f_field if isinstance(h, np.ndarray) else torch.as_tensor(self.f_field, dtype=h.dtype, device=h.device)
        if isinstance(h, np.ndarray):
            f = f.astype(h.dtype)
        fX = 0.25 * (f[1:-1, 1:-1] + f[1:-1, 2:] + f[2:, 1:-1] + f[2:, 2:])
        hX = 0.25 * (h[I] + h[..., 1:-1, 2:] + h[..., 2:, 1:-1] + h[..., 2:, 2:])
        q[I] = (zeta[I] + fX) / hX
        return q, zeta, ke


class MultilayerShallowWater2D(_SWMBase):
    """swm/multilayer.py:95-377."""

    _base_ndim = 3
    _state_cls = MultilayerSW2DState

    def __init__(self, params, consts, grid, strat, modal, f_field, wind_stress_x, wind_stress_y,
                 bc_type="periodic", method="upwind1", dtype="float32", spec=_lib.DEFAULT_SPEC):
        if method != "upwind1":
            raise NotImplementedError('the CUDA path implements method="upwind1" only')
        if bc_type not in ("periodic", "wall"):
            raise ValueError(f"unknown bc_type {bc_type!r}")
        self.params, self.consts, self.grid = params, consts, grid
        self.strat, self.modal = strat, modal
        self.f_field = np.asarray(f_field)
        self.f_field_ml = np.broadcast_to(self.f_field[None], (strat.nl, grid.Ny, grid.Nx))
        self.wind_stress_x, self.wind_stress_y = np.asarray(wind_stress_x), np.asarray(wind_stress_y)
        self.bc_type, self.method = bc_type, method
        self._setup(dtype, strat.nl, grid, bc_type, strat.g_prime, self.f_field, self.wind_stress_x,
                    self.wind_stress_y, np.asarray(strat.H)[0], spec)

    def diagnose(self, state):
        q, zeta, ke = self._diag_fields(state)
        energy, ens, bad = self.diag_scalars(state)
        return MultilayerSW2DDiagnostics(
            energy=energy, total_energy=np.sum(energy, axis=-1), enstrophy=ens,
            total_enstrophy=np.sum(ens, axis=-1), potential_vorticity=q, relative_vorticity=zeta,
            kinetic_energy_field=ke, nonfinite=bad)

    @staticmethod
    def create(nx=64, ny=64, Lx=4e6, Ly=4e6, g=9.81, f0=9.375e-5, beta=1.754e-11, n_layers=3,
               H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125), stratification=None,
               lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
               wind_profile="doublegyre", bc="periodic", method="upwind1", dtype="float32",
               spec=_lib.DEFAULT_SPEC) -> "MultilayerShallowWater2D":
        grid = Grid.from_interior(nx, ny, Lx, Ly)
        if stratification is not None:
            strat = stratification
        else:
            if len(H) != n_layers or len(g_prime) != n_layers:
                raise ValueError(
                    f"n_layers ({n_layers}), len(H) ({len(H)}), and len(g_prime) ({len(g_prime)}) "
                    "must all be equal")
            strat = StratificationProfile.from_layers(H=list(H), g_prime=list(g_prime))
        nl = strat.nl
        modal = ModalTransform.from_stratification(strat, f0)
        params = MultilayerSW2DParams(float(lateral_viscosity), float(bottom_drag), float(wind_amplitude))
        consts = MultilayerSW2DPhysConsts(gravity=g, f0=f0, beta=beta, n_layers=nl)
        f_field, wx, wy = _coriolis_wind(grid, Ly, f0, beta, wind_profile)
        return MultilayerShallowWater2D(params, consts, grid, strat, modal, f_field, wx, wy, bc,
                                        method, dtype, spec)


class NonlinearShallowWater2D(_SWMBase):
    """swm/nonlinear_2d.py:82-318: the nl = 1 case with g_prime = [g]; wind is not divided by a
    layer thickness (H0 = 1) and drag acts on the only layer."""

    _base_ndim = 2
    _state_cls = NonlinearSW2DState

    def __init__(self, params, consts, grid, f_field, wind_stress_x, wind_stress_y,
                 bc_type="periodic", method="upwind1", dtype="float32", spec=_lib.DEFAULT_SPEC):
        if method != "upwind1":
            raise NotImplementedError('the CUDA path implements method="upwind1" only')
        self.params, self.consts, self.grid = params, consts, grid
        self.f_field = np.asarray(f_field)
        self.wind_stress_x, self.wind_stress_y = np.asarray(wind_stress_x), np.asarray(wind_stress_y)
        self.bc_type, self.method = bc_type, method
        self._setup(dtype, 1, grid, bc_type, [consts.gravity], self.f_field, self.wind_stress_x,
                    self.wind_stress_y, 1.0, spec)

    def diagnose(self, state):
        q, zeta, ke = self._diag_fields(state)
        energy, ens, bad = self.diag_scalars(state)
        return NonlinearSW2DDiagnostics(energy=energy[..., 0], enstrophy=ens[..., 0],
                                        potential_vorticity=q, relative_vorticity=zeta,
                                        kinetic_energy_field=ke, nonfinite=bad)

    @staticmethod
    def create(nx=64, ny=64, Lx=1e6, Ly=1e6, g=9.81, f0=1e-4, beta=0.0, H0=100.0,
               lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
               wind_profile="doublegyre", bc="periodic", method="upwind1", dtype="float32",
               spec=_lib.DEFAULT_SPEC) -> "NonlinearShallowWater2D":
        grid = Grid.from_interior(nx, ny, Lx, Ly)
        params = NonlinearSW2DParams(float(lateral_viscosity), float(bottom_drag), float(wind_amplitude))
        consts = NonlinearSW2DPhysConsts(gravity=g, f0=f0, beta=beta, H0=H0)
        f_field, wx, wy = _coriolis_wind(grid, Ly, f0, beta, wind_profile)
        return NonlinearShallowWater2D(params, consts, grid, f_field, wx, wy, bc, method, dtype, spec)


def _coriolis_wind(grid: Grid, Ly, f0, beta, wind_profile):
    """f0 + beta (y - y0) at T points and the normalised wind stress (swm/multilayer.py:345-359)."""
    y = np.arange(grid.Ny, dtype=np.float64) * grid.dy
    Y = np.broadcast_to(y[:, None], (grid.Ny, grid.Nx)).copy()
    f_field = f0 + beta * (Y - Ly / 2.0)
    if wind_profile == "single":
        wx = -np.cos(np.pi * Y / Ly)
    else:
        wx = -np.cos(2.0 * np.pi * Y / Ly)
    return f_field, wx, np.zeros_like(wx)
