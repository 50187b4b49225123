"""Barotropic and baroclinic quasi-geostrophic models on the B200 path.

Mirror of somax/_src/models/qg/baroclinic.py:23-332 and qg/barotropic.py:21-248: same class
names, fields, methods and ``create`` signatures.  ``vector_field``, ``apply_boundary_conditions``,
``_invert_pv``, ``integrate`` and the scalar part of ``diagnose`` call ``libsomax_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib
from ..core import (DeviceIO, Diagnostics, Grid, ModalTransform, Params, PhysConsts, SomaxModel,
                    State, StratificationProfile, scalar, stream_ptr, torch)


# ---------------------------------------------------------------------------- types
@dataclass
class BaroclinicQGState(State):
    """q: layer PV anomaly on T points, (nl, Ny, Nx) (or (members, nl, Ny, Nx))."""
    q: object


@dataclass
class BarotropicQGState(State):
    """q: PV anomaly, (Ny, Nx) (or (members, Ny, Nx))."""
    q: object


@dataclass
class BaroclinicQGParams(Params):
    lateral_viscosity: object
    bottom_drag: object
    wind_amplitude: object


@dataclass
class BarotropicQGParams(Params):
    lateral_viscosity: object
    bottom_drag: object
    wind_amplitude: object


@dataclass(frozen=True)
class BaroclinicQGPhysConsts(PhysConsts):
    f0: float = 1e-4
    beta: float = 1.6e-11
    n_layers: int = 2


@dataclass(frozen=True)
class BarotropicQGPhysConsts(PhysConsts):
    f0: float = 1e-4
    beta: float = 1.6e-11


@dataclass
class BaroclinicQGDiagnostics(Diagnostics):
    psi: object
    u: object
    v: object
    kinetic_energy: object
    total_kinetic_energy: object
    enstrophy: object
    total_enstrophy: object
    relative_vorticity: object
    rossby_radii: object
    nonfinite: object = None


@dataclass
class BarotropicQGDiagnostics(Diagnostics):
    psi: object
    u: object
    v: object
    kinetic_energy: object
    enstrophy: object
    relative_vorticity: object
    nonfinite: object = None


# ---------------------------------------------------------------------------- shared engine
class _QGEngine:
    """Owns the device handles (one per ensemble size) of one QG model."""

    def __init__(self, dtype, nl, ny, nx, dx, dy, Cl2m, Cm2l, lambdas, beta_y, wind, solver, spec):
        self.dtype = np.dtype(dtype)
        self.nl, self.ny, self.nx, self.dx, self.dy = nl, ny, nx, float(dx), float(dy)
        self.Cl2m = np.ascontiguousarray(Cl2m, np.float64)
        self.Cm2l = np.ascontiguousarray(Cm2l, np.float64)
        self.lambdas = np.ascontiguousarray(lambdas, np.float64)
        self.beta_y = np.ascontiguousarray(beta_y, np.float64)
        self.wind = np.ascontiguousarray(wind, np.float64)
        self.solver, self.spec = solver, spec
        self._handles = {}

    def handle(self, batch: int):
        if batch not in self._handles:
            h = C.c_void_p()
            _lib.check(_lib.lib().somax_b200_qg_create(
                C.byref(h), _lib.F32 if self.dtype == np.float32 else _lib.F64, batch, self.nl,
                self.ny, self.nx, self.dx, self.dy, self.Cl2m.ctypes.data, self.Cm2l.ctypes.data,
                self.lambdas.ctypes.data, self.beta_y.ctypes.data, self.wind.ctypes.data,
                self.solver, self.spec))
            self._handles[batch] = h
        return self._handles[batch]

    def close(self):
        for h in self._handles.values():
            _lib.lib().somax_b200_qg_destroy(h)
        self._handles = {}

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def shape4(self, q, base_ndim):
        """(batch, nl, Ny, Nx) view of a caller array with base_ndim dims (+1 with members)."""
        shp = tuple(q.shape)
        core = (self.nl, self.ny + 2, self.nx + 2) if base_ndim == 3 else (self.ny + 2, self.nx + 2)
        if shp == core:
            return 1
        if len(shp) == base_ndim + 1 and shp[1:] == core:
            return shp[0]
        raise ValueError(f"state shape {shp} does not match the model grid {core}")


def _params_struct(params, H0):
    return _lib.ParamsStruct(scalar(params.lateral_viscosity), scalar(params.bottom_drag),
                             scalar(params.wind_amplitude), float(H0))


class _QGBase(SomaxModel):
    _base_ndim = 3
    _state_cls = BaroclinicQGState

    # set by subclasses
    params: object
    grid: Grid
    _engine: _QGEngine
    _H0: float

    def _call_q(self, state, fn):
        io = DeviceIO(self._engine.dtype, self.__dict__.setdefault("_io_cache", {}))
        q = io.to_device(state.q)
        batch = self._engine.shape4(q, self._base_ndim)
        out = fn(self._engine.handle(batch), q, io)
        return out, io

    def _invert_pv(self, q):
        io = DeviceIO(self._engine.dtype, self.__dict__.setdefault("_io_cache", {}))
        qd = io.to_device(q)
        batch = self._engine.shape4(qd, self._base_ndim)
        psi = torch.empty_like(qd)
        _lib.check(_lib.lib().somax_b200_qg_invert(self._engine.handle(batch), qd.data_ptr(),
                                                   psi.data_ptr(), stream_ptr()))
        return io.from_device(psi)

    def vector_field(self, t, state, args=None):
        def run(h, q, io):
            dq = torch.empty_like(q)
            p = _params_struct(self.params, self._H0)
            _lib.check(_lib.lib().somax_b200_qg_rhs(h, q.data_ptr(), dq.data_ptr(), None, C.byref(p),
                                                    0, stream_ptr()))
            return io.from_device(dq)

        out, _ = self._call_q(state, run)
        return self._state_cls(q=out)

    def apply_boundary_conditions(self, state):
        def run(h, q, io):
            out = torch.empty_like(q)
            _lib.check(_lib.lib().somax_b200_qg_apply_bc(h, q.data_ptr(), out.data_ptr(), stream_ptr()))
            return io.from_device(out)

        out, _ = self._call_q(state, run)
        return self._state_cls(q=out)

    def _advance(self, state, n_steps, dt, dt_last, resume=False):
        def run(h, q, io):
            p = _params_struct(self.params, self._H0)
            fn = _lib.lib().somax_b200_qg_resume if resume else _lib.lib().somax_b200_qg_steps
            _lib.check(fn(h, q.data_ptr(), int(n_steps), float(dt), float(dt_last), C.byref(p), stream_ptr()))
            return io.from_device(q)

        out, io = self._call_q(state, run)
        self.last_io = io
        return self._state_cls(q=out)

    def diag_scalars(self, state):
        """(KE[nl], enstrophy[nl], nonfinite) per member from the fused device reduction."""
        def run(h, q, io):
            batch = self._engine.shape4(q, self._base_ndim)
            nl = self._engine.nl
            out = torch.zeros((batch, 2 * nl + 1), dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib().somax_b200_qg_diag(h, q.data_ptr(), out.data_ptr(), stream_ptr()))
            return out.cpu().numpy(), batch

        (vals, batch), _ = self._call_q(state, run)
        nl = self._engine.nl
        ke, ens, bad = vals[:, :nl], vals[:, nl:2 * nl], vals[:, 2 * nl]
        if len(tuple(state.q.shape)) == self._base_ndim:
            ke, ens, bad = ke[0], ens[0], bad[0]
        return ke, ens, bad

    def _diag_fields(self, state):
        """psi, u, v, zeta with the reference's interior-only operator semantics
        (qg/baroclinic.py:199-216); elementwise device ops, not on the hot path."""
        psi = self._invert_pv(state.q)
        xp = np if isinstance(psi, np.ndarray) else torch
        dx, dy = self.grid.dx, self.grid.dy
        u = xp.zeros_like(psi)
        v = xp.zeros_like(psi)
        z = xp.zeros_like(psi)
        c = psi[..., 1:-1, 1:-1]
        u[..., 1:-1, 1:-1] = -((psi[..., 2:, 1:-1] - c) / dy)
        v[..., 1:-1, 1:-1] = (psi[..., 1:-1, 2:] - c) / dx
        z[..., 1:-1, 1:-1] = (psi[..., 1:-1, 2:] - 2 * c + psi[..., 1:-1, :-2]) / (dx * dx) + (
            psi[..., 2:, 1:-1] - 2 * c + psi[..., :-2, 1:-1]) / (dy * dy)
        return psi, u, v, z


# ---------------------------------------------------------------------------- models
class BaroclinicQG(_QGBase):
    """Multilayer QG (qg/baroclinic.py:92-332)."""

    _base_ndim = 3
    _state_cls = BaroclinicQGState

    def __init__(self, params, consts, grid, modal, strat, beta_y, wind_forcing, helmholtz_lambdas,
                 poisson_bc="dst", dtype="float32", solver=_lib.SOLVER_AUTO, spec=_lib.DEFAULT_SPEC):
        if poisson_bc != "dst":
            raise NotImplementedError('the CUDA path implements poisson_bc="dst" only')
        self.params, self.consts, self.grid = params, consts, grid
        self.modal, self.strat = modal, strat
        self.beta_y, self.wind_forcing = np.asarray(beta_y), np.asarray(wind_forcing)
        self.helmholtz_lambdas = np.asarray(helmholtz_lambdas, np.float64)
        self.poisson_bc = poisson_bc
        self.dtype = np.dtype(dtype)
        self._H0 = float(np.asarray(strat.H)[0])
        self._engine = _QGEngine(self.dtype, strat.nl, grid.Ny - 2, grid.Nx - 2, grid.dx, grid.dy,
                                 modal.Cl2m, modal.Cm2l, self.helmholtz_lambdas, self.beta_y,
                                 self.wind_forcing, solver, spec)

    def diagnose(self, state):
        psi, u, v, zeta = self._diag_fields(state)
        ke, ens, bad = self.diag_scalars(state)
        return BaroclinicQGDiagnostics(
            psi=psi, u=u, v=v, kinetic_energy=ke, total_kinetic_energy=np.sum(ke, axis=-1),
            enstrophy=ens, total_enstrophy=np.sum(ens, axis=-1), relative_vorticity=zeta,
            rossby_radii=self.modal.rossby_radii, nonfinite=bad)

    @staticmethod
    def create(nx=64, ny=64, Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11, n_layers=3,
               H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125), stratification=None,
               lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
               wind_profile="doublegyre", poisson_bc="dst", dtype="float32",
               solver=_lib.SOLVER_AUTO, spec=_lib.DEFAULT_SPEC) -> "BaroclinicQG":
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
        helmholtz_lambdas = f0 ** 2 * modal.eigenvalues
        params = BaroclinicQGParams(lateral_viscosity=float(lateral_viscosity),
                                    bottom_drag=float(bottom_drag),
                                    wind_amplitude=float(wind_amplitude))
        consts = BaroclinicQGPhysConsts(f0=f0, beta=beta, n_layers=nl)
        beta_y, wind = _beta_wind(grid, Ly, beta, wind_profile)
        return BaroclinicQG(params, consts, grid, modal, strat, beta_y, wind, helmholtz_lambdas,
                            poisson_bc, dtype, solver, spec)


class BarotropicQG(_QGBase):
    """Single-layer QG (qg/barotropic.py:77-248): the nl = 1 case with Cl2m = Cm2l = [[1]],
    lambda = [0]; wind is not divided by a layer thickness (H0 = 1)."""

    _base_ndim = 2
    _state_cls = BarotropicQGState

    def __init__(self, params, consts, grid, beta_y, wind_forcing, poisson_bc="dst",
                 dtype="float32", solver=_lib.SOLVER_AUTO, spec=_lib.DEFAULT_SPEC):
        if poisson_bc != "dst":
            raise NotImplementedError('the CUDA path implements poisson_bc="dst" only')
        self.params, self.consts, self.grid = params, consts, grid
        self.beta_y, self.wind_forcing = np.asarray(beta_y), np.asarray(wind_forcing)
        self.poisson_bc = poisson_bc
        self.dtype = np.dtype(dtype)
        self._H0 = 1.0
        one = np.ones((1, 1))
        # BarotropicQG._invert_pv (qg/barotropic.py:113-121) does not zero the ring of psi
        self._engine = _QGEngine(self.dtype, 1, grid.Ny - 2, grid.Nx - 2, grid.dx, grid.dy, one, one,
                                 np.zeros(1), self.beta_y, self.wind_forcing, solver,
                                 spec | _lib.SPEC_KEEP_PSI_RING)

    def diagnose(self, state):
        psi, u, v, zeta = self._diag_fields(state)
        ke, ens, bad = self.diag_scalars(state)
        return BarotropicQGDiagnostics(psi=psi, u=u, v=v, kinetic_energy=ke[..., 0],
                                       enstrophy=ens[..., 0], relative_vorticity=zeta, nonfinite=bad)

    @staticmethod
    def create(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, lateral_viscosity=0.0,
               bottom_drag=0.0, wind_amplitude=0.0, wind_profile="doublegyre", dtype="float32",
               solver=_lib.SOLVER_AUTO, spec=_lib.DEFAULT_SPEC) -> "BarotropicQG":
        grid = Grid.from_interior(nx, ny, Lx, Ly)
        params = BarotropicQGParams(lateral_viscosity=float(lateral_viscosity),
                                    bottom_drag=float(bottom_drag),
                                    wind_amplitude=float(wind_amplitude))
        consts = BarotropicQGPhysConsts(f0=f0, beta=beta)
        beta_y, wind = _beta_wind(grid, Ly, beta, wind_profile)
        return BarotropicQG(params, consts, grid, beta_y, wind, "dst", dtype, solver, spec)


def _beta_wind(grid: Grid, Ly, beta, wind_profile):
    """beta*(y - y0) and the normalised wind-stress curl (qg/baroclinic.py:307-318)."""
    y = np.arange(grid.Ny, dtype=np.float64) * grid.dy
    Y = np.broadcast_to(y[:, None], (grid.Ny, grid.Nx)).copy()
    beta_y = beta * (Y - Ly / 2.0)
    if wind_profile == "single":
        wind = np.sin(np.pi * Y / Ly)
    else:
        wind = -np.sin(2.0 * np.pi * Y / Ly)
    return beta_y, wind
