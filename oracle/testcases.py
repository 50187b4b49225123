"""Test-case factories and the synthetic benchmark states.

Test infrastructure only.  ref: somax/_src/models/gfd_testcases.py:184-228 (doublegyre_qg),
:231-287 (doublegyre_baroclinic_qg), :290-362 (baroclinic_instability_swm); parameter values
configs/_authoring/{doublegyre_bt_qg,doublegyre_bc_qg,swm_jet}.py.
"""
from __future__ import annotations

import numpy as np

from . import qg, swm
from .operators import DEFAULT_SPEC


def doublegyre_qg(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, lateral_viscosity=500.0,
                  bottom_drag=1e-7, wind_amplitude=1e-12, spec=DEFAULT_SPEC, dtype=np.float32):
    m = qg.create_barotropic(nx, ny, Lx, Ly, f0, beta, lateral_viscosity, bottom_drag,
                             wind_amplitude, "doublegyre", spec)
    return m, np.zeros((1, ny + 2, nx + 2), dtype)


def doublegyre_baroclinic_qg(nx=128, ny=128, Lx=4e6, Ly=4e6, f0=9.375e-5, beta=1.754e-11,
                             n_layers=3, H=(400.0, 1100.0, 2600.0),
                             g_prime=(9.81, 0.025, 0.0125), lateral_viscosity=15.0,
                             bottom_drag=1e-7, wind_amplitude=1.3e-10, spec=DEFAULT_SPEC,
                             dtype=np.float32):
    m = qg.create_baroclinic(nx, ny, Lx, Ly, f0, beta, n_layers, H, g_prime, lateral_viscosity,
                             bottom_drag, wind_amplitude, "doublegyre", spec)
    return m, np.zeros((n_layers, ny + 2, nx + 2), dtype)


def baroclinic_instability_swm(nx=64, ny=64, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11,
                               H=(500.0, 4500.0), g_prime=(9.81, 0.025),
                               lateral_viscosity=100.0, bottom_drag=1e-7, jet_speed=0.5,
                               jet_width=5e4, perturbation=0.01, spec=DEFAULT_SPEC,
                               dtype=np.float32):
    nl = 2
    m = swm.create_multilayer(nx, ny, Lx, Ly, 9.81, f0, beta, nl, H, g_prime, lateral_viscosity,
                              bottom_drag, 0.0, "doublegyre", "periodic", spec)
    Ny, Nx = ny + 2, nx + 2
    x = np.arange(Nx, dtype=np.float64) * m.dx
    y = np.arange(Ny, dtype=np.float64) * m.dy
    X, Y = np.meshgrid(x, y)
    prof = np.exp(-0.5 * ((Y - Ly / 2.0) / jet_width) ** 2)
    signs = np.array([1.0, -1.0])
    u0 = signs[:, None, None] * jet_speed * prof[None]
    v0 = np.broadcast_to(perturbation * np.sin(4.0 * np.pi * X / Lx)[None] * prof[None],
                         (nl, Ny, Nx)).copy()
    h0 = np.ones((nl, Ny, Nx)) * np.asarray(H, np.float64)[:, None, None]
    return m, (h0.astype(dtype), u0.astype(dtype), v0.astype(dtype))


def synthetic_qg_state(nl, nx, ny, seed=1234, amps=(4e-6, 2e-6, 1e-6), nmodes=8,
                       dtype=np.float32):
    """Seeded low-wavenumber sine superposition: the factory state q0=0
    is degenerate for parity and timing.  Ring = 0."""
    i = np.arange(1, nx + 1, dtype=np.float64)
    j = np.arange(1, ny + 1, dtype=np.float64)
    q = np.zeros((nl, ny + 2, nx + 2))
    mm = np.arange(1, nmodes + 1, dtype=np.float64)
    sx = np.sin(np.pi * mm[:, None] * i[None, :] / (nx + 1))  # (m, i)
    sy = np.sin(np.pi * mm[:, None] * j[None, :] / (ny + 1))  # (n, j)
    w = 1.0 / np.sqrt(mm[:, None] ** 2 + mm[None, :] ** 2)
    for k in range(nl):
        rng = np.random.default_rng(seed + k)
        a = rng.standard_normal((nmodes, nmodes)) * w  # a[m, n]
        q[k, 1:-1, 1:-1] = amps[k % len(amps)] * np.einsum("mn,nj,mi->ji", a, sy, sx)
    return q.astype(dtype)
