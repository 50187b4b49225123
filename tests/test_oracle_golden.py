"""CPU: the oracle reproduces the committed golden fixtures (tools/make_golden.py)."""
from pathlib import Path

import numpy as np

from oracle import qg, swm

G = Path(__file__).parent / "golden"


def test_qg_golden():
    g = np.load(G / "qg3_32x32_f64.npz")
    m = qg.create_baroclinic(nx=32, ny=32, lateral_viscosity=15.0, bottom_drag=1e-7,
                             wind_amplitude=1.3e-10)
    assert np.allclose(m.lambdas, g["lambdas"], rtol=1e-13)
    assert np.allclose(m.invert_pv(g["q0"]), g["psi0"], rtol=0, atol=1e-13 * np.abs(g["psi0"]).max())
    assert np.allclose(m.rhs(m.bc(g["q0"])), g["dq0"], rtol=0, atol=1e-13 * np.abs(g["dq0"]).max())
    q1 = m.integrate(g["q0"], 0.0, float(g["t1"]), float(g["dt"]))
    assert np.allclose(q1, g["q1"], rtol=0, atol=1e-13 * np.abs(g["q1"]).max())
    assert np.allclose(m.diagnose(q1)["kinetic_energy"], g["ke"], rtol=1e-12)


def test_swm_golden():
    g = np.load(G / "swm2_32x32_f64.npz")
    m = swm.create_multilayer(nx=32, ny=32, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, n_layers=2,
                              H=(500.0, 4500.0), g_prime=(9.81, 0.025), lateral_viscosity=100.0,
                              bottom_drag=1e-7, wind_amplitude=1e-6, bc="periodic")
    out = m.integrate(g["h0"], g["u0"], g["v0"], 0.0, float(g["t1"]), float(g["dt"]))
    for a, n in zip(out, "huv"):
        assert np.allclose(a, g[n + "1"], rtol=1e-13, atol=1e-13)
    assert np.allclose(m.diagnose(*out)["energy"], g["energy"], rtol=1e-12)
