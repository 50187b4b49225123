"""Barotropic / baroclinic QG: numpy restatement of the reference models.

Test infrastructure only.  Elliptic conventions pinned by the reference's executed tutorials
(oracle/elliptic.py); see oracle/__init__.py for what is and is not pinned.
ref: somax/_src/models/qg/baroclinic.py:135-228,277-332; qg/barotropic.py:113-248.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import operators as op
from . import tsit5
from .elliptic import helmholtz_dst
from .modal import ModalTransform
from .operators import DEFAULT_SPEC, OperatorSpec


@dataclass
class QGModel:
    """State is ``q`` with shape (nl, Ny, Nx) (barotropic: nl = 1).

    The barotropic model (qg/barotropic.py:123-152) is the nl=1 case with
    Cl2m=Cm2l=[[1]], lambda=[0] and H0=1 (wind not divided by H; drag on the only layer) and
    ``zero_psi_ring=False``: BarotropicQG._invert_pv (qg/barotropic.py:113-121) returns the
    solver's output as it is, while BaroclinicQG._invert_pv zeroes the ring of psi
    (qg/baroclinic.py:157-158).  With the full-array DST solve the ring of psi is not zero.
    """

    nx: int
    ny: int
    dx: float
    dy: float
    Cl2m: np.ndarray
    Cm2l: np.ndarray
    lambdas: np.ndarray
    beta_y: np.ndarray  # (Ny, Nx)
    wind: np.ndarray  # (Ny, Nx)
    H0: float
    nu: float
    kappa: float
    tau0: float
    rossby_radii: np.ndarray | None = None
    spec: OperatorSpec = field(default_factory=lambda: DEFAULT_SPEC)
    workers: int | None = None
    zero_psi_ring: bool = True

    @property
    def nl(self):
        return self.Cl2m.shape[0]

    # ref: qg/baroclinic.py:192-195, qg/barotropic.py:154-157
    def bc(self, q):
        return op.zero_boundaries(q)

    # ref: qg/baroclinic.py:135-159
    def invert_pv(self, q):
        dt = q.dtype
        qm = np.einsum("lm,m...->l...", self.Cl2m.astype(dt), q)
        pm = helmholtz_dst(qm, self.dx, self.dy, self.lambdas, self.spec, self.workers)
        psi = np.einsum("lm,m...->l...", self.Cm2l.astype(dt), pm)
        return op.zero_boundaries(psi) if self.zero_psi_ring else psi

    # ref: qg/baroclinic.py:161-190
    def rhs(self, q):
        dt = q.dtype
        psi = self.invert_pv(q)
        q_total = q + self.beta_y.astype(dt)[None]
        J = op.arakawa_jacobian(psi, q_total, self.dx, self.dy)
        dq = np.zeros_like(q)
        dq[:, 1:-1, 1:-1] = -J
        dq[0] += (dt.type(self.tau0) * self.wind.astype(dt)) / dt.type(self.H0)
        dq[-1] += -self.kappa * op.laplacian(psi[-1], self.dx, self.dy)
        dq = dq + self.nu * op.laplacian(q, self.dx, self.dy)
        return dq

    def integrate(self, q0, t0, t1, dt, on_step=None):
        (q,) = tsit5.integrate(
            lambda y: (self.rhs(y[0]),), lambda y: (self.bc(y[0]),), (q0,), t0, t1, dt,
            on_step=(lambda i, y: on_step(i, y[0])) if on_step else None,
        )
        return q

    # ref: qg/baroclinic.py:197-228, qg/barotropic.py:159-183
    def diagnose(self, q):
        psi = self.invert_pv(q)
        u = -op.diff_y_T_to_V(psi, self.dy)
        v = op.diff_x_T_to_U(psi, self.dx)
        s = (slice(None), slice(1, -1), slice(1, -1))
        uT, vT = op.V_to_T(u), op.U_to_T(v)
        area = self.dx * self.dy
        ke = 0.5 * np.sum(uT[s] ** 2 + vT[s] ** 2, axis=(-2, -1), dtype=np.float64) * area
        ens = 0.5 * np.sum(q[s] ** 2, axis=(-2, -1), dtype=np.float64) * area
        zeta = op.laplacian(psi, self.dx, self.dy)
        return dict(psi=psi, u=u, v=v, kinetic_energy=ke, total_kinetic_energy=ke.sum(),
                    enstrophy=ens, total_enstrophy=ens.sum(), relative_vorticity=zeta,
                    rossby_radii=self.rossby_radii)


def _yfields(nx, ny, Lx, Ly):
    Ny, Nx = ny + 2, nx + 2
    dy = Ly / ny
    y = np.arange(Ny, dtype=np.float64) * dy
    return np.broadcast_to(y[:, None], (Ny, Nx)).copy()


def _wind(Y, Ly, profile):
    # ref: qg/baroclinic.py:313-318
    if profile == "single":
        return np.sin(np.pi * Y / Ly)
    return -np.sin(2.0 * np.pi * Y / Ly)


def create_baroclinic(nx=64, ny=64, Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11, n_layers=3,
                      H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125),
                      lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
                      wind_profile="doublegyre", spec=DEFAULT_SPEC) -> QGModel:
    """ref: BaroclinicQG.create, qg/baroclinic.py:230-332."""
    if len(H) != n_layers or len(g_prime) != n_layers:
        raise ValueError(
            f"n_layers ({n_layers}), len(H) ({len(H)}), and len(g_prime) ({len(g_prime)}) "
            "must all be equal")
    modal = ModalTransform.from_physics(H, g_prime, f0)
    Y = _yfields(nx, ny, Lx, Ly)
    return QGModel(nx=nx, ny=ny, dx=Lx / nx, dy=Ly / ny, Cl2m=modal.Cl2m, Cm2l=modal.Cm2l,
                   lambdas=f0 ** 2 * modal.eigenvalues, beta_y=beta * (Y - Ly / 2.0),
                   wind=_wind(Y, Ly, wind_profile), H0=float(H[0]), nu=lateral_viscosity,
                   kappa=bottom_drag, tau0=wind_amplitude, rossby_radii=modal.rossby_radii,
                   spec=spec)


def create_barotropic(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11,
                      lateral_viscosity=0.0, bottom_drag=0.0, wind_amplitude=0.0,
                      wind_profile="doublegyre", spec=DEFAULT_SPEC) -> QGModel:
    """ref: BarotropicQG.create, qg/barotropic.py:185-248."""
    Y = _yfields(nx, ny, Lx, Ly)
    one = np.ones((1, 1))
    return QGModel(nx=nx, ny=ny, dx=Lx / nx, dy=Ly / ny, Cl2m=one, Cm2l=one.copy(),
                   lambdas=np.zeros(1), beta_y=beta * (Y - Ly / 2.0),
                   wind=_wind(Y, Ly, wind_profile), H0=1.0, nu=lateral_viscosity,
                   kappa=bottom_drag, tau0=wind_amplitude, spec=spec, zero_psi_ring=False)
