"""Multilayer / single-layer nonlinear shallow water (vector-invariant form).

Test infrastructure only.  Operator conventions pinned by the reference's printed tutorial outputs
(step17_shallow_water_2d.ipynb; tests/test_oracle_reference_pins.py; see oracle/__init__.py).
ref: somax/_src/models/swm/multilayer.py:150-256,313-410; swm/nonlinear_2d.py:132-234.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import operators as op
from . import tsit5
from .operators import DEFAULT_SPEC, OperatorSpec


@dataclass
class SWMModel:
    """State is (h, u, v), each (nl, Ny, Nx).  NonlinearShallowWater2D
    (swm/nonlinear_2d.py:132-181) is nl=1 with g_prime=[g], H0=1."""

    nx: int
    ny: int
    dx: float
    dy: float
    g_prime: np.ndarray  # (nl,)
    f_field: np.ndarray  # (Ny, Nx), T points
    wind_x: np.ndarray
    wind_y: np.ndarray
    H0: float
    nu: float
    kappa: float
    tau0: float
    bc_type: str = "periodic"
    H: np.ndarray | None = None
    spec: OperatorSpec = field(default_factory=lambda: DEFAULT_SPEC)

    @property
    def nl(self):
        return len(self.g_prime)

    # ref: swm/multilayer.py:203-223
    def bc(self, h, u, v):
        if self.bc_type == "periodic":
            return op.enforce_periodic(h), op.enforce_periodic(u), op.enforce_periodic(v)
        return op.wall_bc_h(h), op.wall_bc_u(u), op.wall_bc_v(v)

    # ref: swm/multilayer.py:150-201
    def rhs(self, h, u, v):
        dt = h.dtype
        dx, dy, sp = self.dx, self.dy, self.spec
        dh = op.advection_upwind1(h, u, v, dx, dy, sp)
        q = op.potential_vorticity(u, v, h, self.f_field.astype(dt)[None], dx, dy)
        uh = op.T_to_U(h) * u
        vh = op.T_to_V(h) * v
        qU, qV = op.X_to_U(q), op.X_to_V(q)
        vhU, uhV = op.V_to_U(vh), op.U_to_V(uh)
        ke = op.kinetic_energy(u, v)
        p = np.cumsum(self.g_prime.astype(dt)[:, None, None] * h, axis=0, dtype=dt)
        P = ke + p
        du = qU * vhU - op.diff_x_T_to_U(P, dx)
        dv = -qV * uhV - op.diff_y_T_to_V(P, dy)
        du[0] += (dt.type(self.tau0) * self.wind_x.astype(dt)) / dt.type(self.H0)
        dv[0] += (dt.type(self.tau0) * self.wind_y.astype(dt)) / dt.type(self.H0)
        du = du + op.diffusion(u, self.nu, dx, dy, sp)
        dv = dv + op.diffusion(v, self.nu, dx, dy, sp)
        du[-1] += -self.kappa * u[-1]
        dv[-1] += -self.kappa * v[-1]
        return dh, du, dv

    def integrate(self, h0, u0, v0, t0, t1, dt, on_step=None):
        return tsit5.integrate(lambda y: self.rhs(*y), lambda y: self.bc(*y), (h0, u0, v0),
                               t0, t1, dt, on_step=on_step)

    # ref: swm/multilayer.py:225-256
    def diagnose(self, h, u, v):
        dt = h.dtype
        s = (slice(None), slice(1, -1), slice(1, -1))
        ke = op.kinetic_energy(u, v)
        zeta = op.curl(u, v, self.dx, self.dy)
        q = op.potential_vorticity(u, v, h, self.f_field.astype(dt)[None], self.dx, self.dy)
        hX = op.T_to_X(h)
        area = self.dx * self.dy
        f8 = np.float64
        ke_sum = np.sum(ke[s], axis=(-2, -1), dtype=f8) * area
        pe_sum = 0.5 * self.g_prime * np.sum(h[s] ** 2, axis=(-2, -1), dtype=f8) * area
        energy = ke_sum + pe_sum
        ens = 0.5 * np.sum(q[s] ** 2 * hX[s], axis=(-2, -1), dtype=f8) * area
        return dict(energy=energy, total_energy=energy.sum(), enstrophy=ens,
                    total_enstrophy=ens.sum(), potential_vorticity=q, relative_vorticity=zeta,
                    kinetic_energy_field=ke)


def create_multilayer(nx=64, ny=64, Lx=4e6, Ly=4e6, g=9.81, f0=9.375e-5, beta=1.754e-11,
                      n_layers=3, H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125),
                      lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
                      wind_profile="doublegyre", bc="periodic", spec=DEFAULT_SPEC) -> SWMModel:
    """ref: MultilayerShallowWater2D.create, swm/multilayer.py:258-377."""
    if len(H) != n_layers or len(g_prime) != n_layers:
        raise ValueError(
            f"n_layers ({n_layers}), len(H) ({len(H)}), and len(g_prime) ({len(g_prime)}) "
            "must all be equal")
    Ny, Nx = ny + 2, nx + 2
    dy = Ly / ny
    y = np.arange(Ny, dtype=np.float64) * dy
    Y = np.broadcast_to(y[:, None], (Ny, Nx)).copy()
    f_field = f0 + beta * (Y - Ly / 2.0)
    if wind_profile == "single":
        wx = -np.cos(np.pi * Y / Ly)
    else:
        wx = -np.cos(2.0 * np.pi * Y / Ly)
    return SWMModel(nx=nx, ny=ny, dx=Lx / nx, dy=dy, g_prime=np.asarray(g_prime, np.float64),
                    f_field=f_field, wind_x=wx, wind_y=np.zeros_like(wx), H0=float(H[0]),
                    nu=lateral_viscosity, kappa=bottom_drag, tau0=wind_amplitude, bc_type=bc,
                    H=np.asarray(H, np.float64), spec=spec)
