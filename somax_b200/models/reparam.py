"""Reparameterized QG model: multilayer shallow water + geostrophic projection.

Host-side mirror of ``somax/_src/models/qg/reparameterized.py:66-330`` (same class, field, method
and factory names).  The state is the shallow-water (h, u, v); ``apply_boundary_conditions``
applies the shallow-water wall BCs and then projects onto the geostrophic manifold,
``P = G (Q G)^-1 Q``.  On the device that is two small stencil kernels around the PV-inversion
solver of the QG models, enabled on a shallow-water handle by ``somax_b200_swm_set_projection``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib
from ..core import Diagnostics, SomaxModel, stream_ptr
from .swm import MultilayerShallowWater2D, MultilayerSW2DState

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


@dataclass
class ReparamQGDiagnostics(Diagnostics):
    """qg/reparameterized.py:37-63."""

    energy: object
    total_energy: object
    enstrophy: object
    total_enstrophy: object
    potential_vorticity: object
    relative_vorticity: object
    kinetic_energy_field: object
    psi: object
    u_ageostrophic: object
    v_ageostrophic: object
    nonfinite: object = None


class _ProjectedSWM(MultilayerShallowWater2D):
    """The shallow-water engine whose device handles carry the projection."""

    def _handle(self, batch):
        fresh = batch not in self._handles
        h = super()._handle(batch)
        if fresh:
            m, s = self.modal, self.strat
            arr = [np.ascontiguousarray(a, np.float64) for a in
                   (s.H, m.Cl2m, m.Cm2l, m.eigenvalues, self._helmholtz_lambdas)]
            self._proj_keepalive = arr
            _lib.check(_lib.lib().somax_b200_swm_set_projection(
                h, float(self.consts.f0), *[a.ctypes.data for a in arr], self._solver))
        return h


class ReparameterizedQG(SomaxModel):
    """qg/reparameterized.py:66-330.  ``swm`` is the plain shallow-water model (its own BCs and
    vector field, as in the reference); the projection lives in this class's methods."""

    def __init__(self, swm: MultilayerShallowWater2D, helmholtz_lambdas, poisson_bc="dst",
                 solver=_lib.SOLVER_AUTO):
        if poisson_bc != "dst":
            raise NotImplementedError('the CUDA path implements poisson_bc="dst" only')
        self.swm = swm
        self.helmholtz_lambdas = np.asarray(helmholtz_lambdas, np.float64)
        self.poisson_bc = poisson_bc
        eng = _ProjectedSWM(swm.params, swm.consts, swm.grid, swm.strat, swm.modal, swm.f_field,
                            swm.wind_stress_x, swm.wind_stress_y, swm.bc_type, swm.method,
                            swm.dtype.name, swm._spec)
        eng._helmholtz_lambdas, eng._solver = self.helmholtz_lambdas, solver
        self._eng = eng

    # --- delegated properties (reparameterized.py:95-126) ---
    params = property(lambda self: self.swm.params)
    consts = property(lambda self: self.swm.consts)
    grid = property(lambda self: self.swm.grid)
    strat = property(lambda self: self.swm.strat)
    modal = property(lambda self: self.swm.modal)
    dtype = property(lambda self: self.swm.dtype)

    @property
    def last_io(self):
        return self._eng.last_io

    def project(self, state):
        """P = G (Q G)^-1 Q on the state as given (reparameterized.py:142-177)."""
        io, h, u, v, hd = self._eng._dev(state)
        ho, uo, vo = torch.empty_like(h), torch.empty_like(u), torch.empty_like(v)
        _lib.check(_lib.lib().somax_b200_swm_project(hd, h.data_ptr(), u.data_ptr(), v.data_ptr(),
                                                     ho.data_ptr(), uo.data_ptr(), vo.data_ptr(), stream_ptr()))
        return MultilayerSW2DState(h=io.from_device(ho), u=io.from_device(uo), v=io.from_device(vo))

    def vector_field(self, t, state, args=None):
        """The shallow-water vector field (reparameterized.py:179-183)."""
        return self._eng.vector_field(t, state, args)

    def apply_boundary_conditions(self, state):
        """Shallow-water BCs, then the projection (reparameterized.py:185-188)."""
        return self._eng.apply_boundary_conditions(state)

    def _advance(self, state, n_steps, dt, dt_last, resume=False):
        return self._eng._advance(state, n_steps, dt, dt_last, resume=resume)

    def _solve_helmholtz(self, q):
        """(Q G)^-1: PV -> pressure / streamfunction via the modal Helmholtz solve, ring zeroed
        (reparameterized.py:128-140) - the QG models' PV inversion with this model's lambdas."""
        from .qg import BaroclinicQG, BaroclinicQGParams, BaroclinicQGPhysConsts
        if getattr(self, "_inv", None) is None:
            zero = np.zeros((self.grid.Ny, self.grid.Nx))
            self._inv = BaroclinicQG(BaroclinicQGParams(0.0, 0.0, 0.0),
                                     BaroclinicQGPhysConsts(f0=self.consts.f0, beta=self.consts.beta,
                                                            n_layers=self.strat.nl),
                                     self.grid, self.modal, self.strat, zero, zero, self.helmholtz_lambdas,
                                     dtype=self.dtype.name, solver=self._eng._solver)
        return self._inv._invert_pv(q)

    def diagnose(self, state):
        """reparameterized.py:190-226: the shallow-water diagnostics plus psi and the ageostrophic
        velocity u - u_g, v - v_g (grad_perp of psi); elementwise host / device ops, not hot."""
        d = self.swm.diagnose(state)
        h, u, v = state.h, state.u, state.v
        xp = np if isinstance(h, np.ndarray) else torch
        f0, dx, dy = self.consts.f0, self.grid.dx, self.grid.dy
        H = np.asarray(self.strat.H, np.float64)
        Hb = H[:, None, None] if isinstance(h, np.ndarray) else torch.as_tensor(H, dtype=h.dtype, device=h.device)[:, None, None]
        if isinstance(h, np.ndarray):
            Hb = Hb.astype(h.dtype)
        q = d.relative_vorticity - f0 * (h - Hb) / Hb
        psi = self._solve_helmholtz(q)
        ug, vg = xp.zeros_like(psi), xp.zeros_like(psi)
        ug[..., 1:-1, 1:-1] = -((psi[..., 1:-1, 1:-1] - psi[..., :-2, 1:-1]) / dy)
        vg[..., 1:-1, 1:-1] = (psi[..., 1:-1, 1:-1] - psi[..., 1:-1, :-2]) / dx
        return ReparamQGDiagnostics(
            energy=d.energy, total_energy=d.total_energy, enstrophy=d.enstrophy,
            total_enstrophy=d.total_enstrophy, potential_vorticity=d.potential_vorticity,
            relative_vorticity=d.relative_vorticity, kinetic_energy_field=d.kinetic_energy_field,
            psi=psi, u_ageostrophic=u - ug, v_ageostrophic=v - vg, nonfinite=d.nonfinite)

    def diag_scalars(self, state):
        return self.swm.diag_scalars(state)

    @staticmethod
    def create(nx=64, ny=64, Lx=4e6, Ly=4e6, g=9.81, f0=9.375e-5, beta=1.754e-11, n_layers=3,
               H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125), stratification=None,
               lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0, wind_profile="doublegyre",
               bc="wall", method="upwind1", poisson_bc="dst", dtype="float32", solver=_lib.SOLVER_AUTO,
               spec=_lib.DEFAULT_SPEC) -> "ReparameterizedQG":
        if bc != "wall":
            raise ValueError(
                f"ReparameterizedQG requires wall BCs (got bc={bc!r}). The geostrophic projection uses "
                "Dirichlet Helmholtz inversion which is incompatible with periodic BCs.")
        swm = MultilayerShallowWater2D.create(
            nx=nx, ny=ny, Lx=Lx, Ly=Ly, g=g, f0=f0, beta=beta, n_layers=n_layers, H=H, g_prime=g_prime,
            stratification=stratification, lateral_viscosity=lateral_viscosity, bottom_drag=bottom_drag,
            wind_amplitude=wind_amplitude, wind_profile=wind_profile, bc=bc, method=method, dtype=dtype, spec=spec)
        return ReparameterizedQG(swm, f0 ** 2 * swm.modal.eigenvalues, poisson_bc, solver)
