"""Pre-configured GFD test cases on the B200 path: ``(model, state0)`` factories.

Mirror of somax/_src/models/gfd_testcases.py for the four ``somax-sim`` test cases
(cli/_factories.py:153-158): barotropic_jet_instability (:120-181), doublegyre_qg (:184-228),
doublegyre_baroclinic_qg (:231-287), baroclinic_instability_swm (:290-362), and
doublegyre_reparameterized_qg (:365-428).  States are numpy
arrays in the model dtype (pass them, or CUDA tensors, to ``model.integrate``).
"""
from __future__ import annotations

import numpy as np

from .models.qg import BaroclinicQG, BaroclinicQGState, BarotropicQG, BarotropicQGState
from .models.reparam import ReparameterizedQG
from .models.swm import (MultilayerShallowWater2D, MultilayerSW2DState, NonlinearShallowWater2D,
                         NonlinearSW2DState)


def barotropic_jet_instability(nx=128, ny=128, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, H0=1000.0,
                               jet_speed=1.0, jet_width=5e4, perturbation=0.01,
                               lateral_viscosity=100.0, dtype="float32"):
    model = NonlinearShallowWater2D.create(nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=f0, beta=beta, H0=H0,
                                           lateral_viscosity=lateral_viscosity, bc="periodic",
                                           dtype=dtype)
    g = model.grid
    X, Y = np.meshgrid(np.arange(g.Nx) * g.dx, np.arange(g.Ny) * g.dy)
    prof = np.exp(-0.5 * ((Y - Ly / 2.0) / jet_width) ** 2)
    u0 = jet_speed * prof
    v0 = perturbation * np.sin(4.0 * np.pi * X / Lx) * prof
    h0 = np.full_like(u0, H0)
    dt = np.dtype(dtype)
    return model, NonlinearSW2DState(h=h0.astype(dt), u=u0.astype(dt), v=v0.astype(dt))


def doublegyre_qg(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, lateral_viscosity=500.0,
                  bottom_drag=1e-7, wind_amplitude=1e-12, dtype="float32"):
    model = BarotropicQG.create(nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=f0, beta=beta,
                                lateral_viscosity=lateral_viscosity, bottom_drag=bottom_drag,
                                wind_amplitude=wind_amplitude, wind_profile="doublegyre",
                                dtype=dtype)
    q0 = np.zeros((model.grid.Ny, model.grid.Nx), np.dtype(dtype))
    return model, BarotropicQGState(q=q0)


def doublegyre_baroclinic_qg(nx=128, ny=128, Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11,
                             n_layers=3, H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125),
                             lateral_viscosity=15.0, bottom_drag=1e-7, wind_amplitude=1.3e-10,
                             dtype="float32"):
    model = BaroclinicQG.create(nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=f0, beta=beta, n_layers=n_layers,
                                H=H, g_prime=g_prime, lateral_viscosity=lateral_viscosity,
                                bottom_drag=bottom_drag, wind_amplitude=wind_amplitude,
                                wind_profile="doublegyre", dtype=dtype)
    q0 = np.zeros((model.consts.n_layers, model.grid.Ny, model.grid.Nx), np.dtype(dtype))
    return model, BaroclinicQGState(q=q0)


def baroclinic_instability_swm(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11,
                               H=(500.0, 4500.0), g_prime=(9.81, 0.025), lateral_viscosity=100.0,
                               bottom_drag=1e-7, jet_speed=0.5, jet_width=5e4, perturbation=0.01,
                               dtype="float32"):
    nl = 2
    model = MultilayerShallowWater2D.create(nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=f0, beta=beta,
                                            n_layers=nl, H=H, g_prime=g_prime,
                                            lateral_viscosity=lateral_viscosity,
                                            bottom_drag=bottom_drag, bc="periodic", dtype=dtype)
    g = model.grid
    X, Y = np.meshgrid(np.arange(g.Nx) * g.dx, np.arange(g.Ny) * g.dy)
    prof = np.exp(-0.5 * ((Y - Ly / 2.0) / jet_width) ** 2)
    signs = np.array([1.0, -1.0])[:nl]
    u0 = signs[:, None, None] * jet_speed * prof[None]
    v0 = np.broadcast_to(perturbation * np.sin(4.0 * np.pi * X / Lx)[None] * prof[None],
                         (nl, g.Ny, g.Nx)).copy()
    h0 = np.ones((nl, g.Ny, g.Nx)) * np.asarray(model.strat.H)[:, None, None]
    dt = np.dtype(dtype)
    return model, MultilayerSW2DState(h=h0.astype(dt), u=u0.astype(dt), v=v0.astype(dt))


def doublegyre_reparameterized_qg(nx=128, ny=128, Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11, n_layers=3,
                                  H=(400.0, 1100.0, 2600.0), g_prime=(9.81, 0.025, 0.0125),
                                  lateral_viscosity=15.0, bottom_drag=3.6e-8, wind_amplitude=8e-5,
                                  dtype="float32"):
    """Wind-driven multilayer double gyre solved in (u, v, h) with the geostrophic projection
    (gfd_testcases.py:365-428): state at rest, h = H."""
    model = ReparameterizedQG.create(nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=f0, beta=beta, n_layers=n_layers, H=H,
                                     g_prime=g_prime, lateral_viscosity=lateral_viscosity,
                                     bottom_drag=bottom_drag, wind_amplitude=wind_amplitude,
                                     wind_profile="doublegyre", bc="wall", dtype=dtype)
    nl, g = model.consts.n_layers, model.grid
    dt = np.dtype(dtype)
    h0 = (np.ones((nl, g.Ny, g.Nx)) * np.asarray(model.strat.H)[:, None, None]).astype(dt)
    z = np.zeros((nl, g.Ny, g.Nx), dt)
    return model, MultilayerSW2DState(h=h0, u=z, v=z.copy())


def synthetic_qg_state(nl, nx, ny, seed=1234, amps=(4e-6, 2e-6, 1e-6), nmodes=8, dtype="float32"):
    """Seeded low-wavenumber sine superposition used by bench.py and the parity tests
    (SURVEY section 8d): the factory state q0 = 0 is degenerate.  Ghost ring = 0."""
    i = np.arange(1, nx + 1, dtype=np.float64)
    j = np.arange(1, ny + 1, dtype=np.float64)
    mm = np.arange(1, nmodes + 1, dtype=np.float64)
    sx = np.sin(np.pi * mm[:, None] * i[None, :] / (nx + 1))
    sy = np.sin(np.pi * mm[:, None] * j[None, :] / (ny + 1))
    w = 1.0 / np.sqrt(mm[:, None] ** 2 + mm[None, :] ** 2)
    q = np.zeros((nl, ny + 2, nx + 2), np.dtype(dtype))
    for k in range(nl):
        a = np.random.default_rng(seed + k).standard_normal((nmodes, nmodes)) * w
        q[k, 1:-1, 1:-1] = (amps[k % len(amps)] * ((sy.T @ a.T) @ sx)).astype(q.dtype)
    return q


TEST_CASES = {
    "barotropic_jet_instability": barotropic_jet_instability,
    "doublegyre_qg": doublegyre_qg,
    "doublegyre_baroclinic_qg": doublegyre_baroclinic_qg,
    "baroclinic_instability_swm": baroclinic_instability_swm,
    "doublegyre_reparameterized_qg": doublegyre_reparameterized_qg,
}
