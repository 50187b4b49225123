"""Reparameterized QG (multilayer shallow water + geostrophic projection): numpy restatement.

Test infrastructure only.  ref: somax/_src/models/qg/reparameterized.py:128-189 (``_solve_helmholtz``,
``project``, ``vector_field``, ``apply_boundary_conditions``), :190-226 (``diagnose``), :228-330
(``create``).  The shallow-water part and the Helmholtz solve are the pinned pieces of oracle/swm.py
and oracle/elliptic.py.  PARITY UNPINNED for one operator: ``Difference2D.grad_perp`` of finitevolx
(not under /root/reference; no tutorial output exercises it).  It is restated from its call site's
comment - "returns (u@U, v@V) = (-dpsi/dy, dpsi/dx)" (reparameterized.py:169-170) - with psi on the
grid of the relative vorticity it is inverted from (X points): backward differences to the U / V
points, interior only, zero ring (the finitevolx convention every pinned operator follows).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import operators as op
from . import tsit5
from .elliptic import helmholtz_dst
from .modal import ModalTransform
from .operators import DEFAULT_SPEC
from .swm import SWMModel, create_multilayer


def grad_perp(psi, dx, dy):
    """(-d psi/dy at U points, d psi/dx at V points), interior only."""
    u, v = np.zeros_like(psi), np.zeros_like(psi)
    c = psi[..., 1:-1, 1:-1]
    u[..., 1:-1, 1:-1] = -((c - psi[..., :-2, 1:-1]) / dy)
    v[..., 1:-1, 1:-1] = (c - psi[..., 1:-1, :-2]) / dx
    return u, v


@dataclass
class ReparamQGModel:
    swm: SWMModel
    modal: ModalTransform
    lambdas: np.ndarray      # helmholtz_lambdas = f0^2 * eigenvalues
    f0: float
    workers: int | None = None

    # ref: reparameterized.py:128-140
    def solve_helmholtz(self, q):
        pm = helmholtz_dst(self.modal.to_modal(q), self.swm.dx, self.swm.dy, self.lambdas, self.swm.spec, self.workers)
        return op.zero_boundaries(self.modal.to_layer(pm))

    def pv(self, h, u, v):
        dt = h.dtype
        H = self.swm.H.astype(dt)[:, None, None]
        return op.curl(u, v, self.swm.dx, self.swm.dy) - dt.type(self.f0) * (h - H) / H

    # ref: reparameterized.py:142-177
    def project(self, h, u, v):
        dt = h.dtype
        H = self.swm.H.astype(dt)[:, None, None]
        psi = self.solve_helmholtz(self.pv(h, u, v))
        ug, vg = grad_perp(psi, self.swm.dx, self.swm.dy)
        pm = self.modal.to_modal(psi)
        A_psi = self.modal.to_layer(self.modal.eigenvalues.astype(dt)[:, None, None] * pm)
        return H * (1.0 + dt.type(self.f0) * A_psi), ug, vg

    # ref: reparameterized.py:179-188
    def rhs(self, h, u, v):
        return self.swm.rhs(h, u, v)

    def bc(self, h, u, v):
        return self.project(*self.swm.bc(h, u, v))

    def integrate(self, h0, u0, v0, t0, t1, dt, on_step=None):
        return tsit5.integrate(lambda y: self.rhs(*y), lambda y: self.bc(*y), (h0, u0, v0), t0, t1, dt,
                               on_step=on_step)

    # ref: reparameterized.py:190-226
    def diagnose(self, h, u, v):
        d = self.swm.diagnose(h, u, v)
        psi = self.solve_helmholtz(self.pv(h, u, v))
        ug, vg = grad_perp(psi, self.swm.dx, self.swm.dy)
        d.update(psi=psi, u_ageostrophic=u - ug, v_ageostrophic=v - vg)
        return d


def create_reparameterized(nx=64, ny=64, Lx=4e6, Ly=4e6, g=9.81, f0=9.375e-5, beta=1.754e-11, n_layers=3,
                           H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125), lateral_viscosity=0.0,
                           bottom_drag=0.0, wind_amplitude=0.0, wind_profile="doublegyre", bc="wall",
                           spec=DEFAULT_SPEC) -> ReparamQGModel:
    """ref: ReparameterizedQG.create, reparameterized.py:228-330."""
    if bc != "wall":
        raise ValueError(f"ReparameterizedQG requires wall BCs (got bc={bc!r}).")
    swm = create_multilayer(nx=nx, ny=ny, Lx=Lx, Ly=Ly, g=g, f0=f0, beta=beta, n_layers=n_layers, H=H,
                            g_prime=g_prime, lateral_viscosity=lateral_viscosity, bottom_drag=bottom_drag,
                            wind_amplitude=wind_amplitude, wind_profile=wind_profile, bc=bc, spec=spec)
    modal = ModalTransform.from_physics(H, g_prime, f0)
    return ReparamQGModel(swm=swm, modal=modal, lambdas=f0 ** 2 * modal.eigenvalues, f0=f0)
