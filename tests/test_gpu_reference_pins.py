"""The CUDA path against numbers the REFERENCE printed in its executed tutorials
(tests/golden/reference_notebook_outputs.json; see tests/test_oracle_reference_pins.py for what
they are).  No oracle in between: these compare libsomax_b200 with the reference directly.

  * BarotropicQG._invert_pv is ``streamfunction_from_vorticity(q, dx, dy, bc="dst")``
    (qg/barotropic.py:119-121), the very call PoissonSolver2D.solve makes (pde2d/poisson.py:36-38);
  * a one-layer BaroclinicQG with helmholtz_lambdas = [lambda] is HelmholtzSolver2D.solve
    (pde2d/poisson.py:129-133) up to the zeroed ring of psi;
  * NonlinearShallowWater2D through 56 031 Tsit5 steps (step17_shallow_water_2d.ipynb).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_notebook_outputs.json")))["parsed"]


def _grid(n):
    x = np.arange(n + 2) * (1.0 / n)
    return np.meshgrid(x, x)


@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_poisson_errors_match_reference_printout(dtype, solver):
    import somax_b200 as sb
    for n, ref in REF["poisson_dirichlet_l2_by_n"].items():
        n = int(n)
        X, Y = _grid(n)
        exact = np.sin(np.pi * X) * np.sin(np.pi * Y)
        m = sb.BarotropicQG.create(nx=n, ny=n, Lx=1.0, Ly=1.0, dtype=dtype, solver=solver)
        phi = m._invert_pv((-2.0 * np.pi ** 2 * exact).astype(dtype))
        err = (phi.astype(np.float64) - exact)[1:-1, 1:-1]
        assert float(np.sqrt(np.mean(err ** 2))) == pytest.approx(ref, rel=2e-5), n
        if n == 64:
            assert float(np.abs(err).max()) == pytest.approx(REF["poisson_dirichlet_linf_n64"], rel=2e-5)
            ex23 = np.sin(2 * np.pi * X) * np.sin(3 * np.pi * Y)
            phi = m._invert_pv((-13.0 * np.pi ** 2 * ex23).astype(dtype))
            e23 = (phi.astype(np.float64) - ex23)[1:-1, 1:-1]
            assert float(np.sqrt(np.mean(e23 ** 2))) == pytest.approx(REF["poisson_dirichlet_mode23_l2_n64"], rel=2e-5)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_helmholtz_maxima_match_reference_printout(dtype):
    import somax_b200 as sb
    from somax_b200.core import Grid, ModalTransform, StratificationProfile
    from somax_b200.models.qg import BaroclinicQGParams, BaroclinicQGPhysConsts
    n = 64
    X, Y = _grid(n)
    rhs = (np.sin(np.pi * X) * np.sin(np.pi * Y)).astype(dtype)[None]
    grid = Grid.from_interior(n, n, 1.0, 1.0)
    strat = StratificationProfile.from_layers(H=[1.0], g_prime=[1.0])
    one = np.ones((1, 1))
    modal = ModalTransform(Cl2m=one, Cm2l=one, eigenvalues=np.zeros(1), rossby_radii=np.ones(1))
    zero = np.zeros((n + 2, n + 2))
    for lam, ref in REF["helmholtz_maxabs_by_lambda"].items():
        m = sb.BaroclinicQG(BaroclinicQGParams(0.0, 0.0, 0.0), BaroclinicQGPhysConsts(f0=1.0, beta=0.0, n_layers=1),
                            grid, modal, strat, zero, zero, np.array([float(lam)]), dtype=dtype)
        phi = m._invert_pv(rhs)
        assert float(np.abs(phi[0, 1:-1, 1:-1]).max()) == pytest.approx(ref, abs=1e-6), lam


def test_swm_spinup_matches_reference_printout():
    """step17_shallow_water_2d.ipynb cells 3-9: wind-driven spin-up of a 32^2 closed basin to
    t = 5e6 s in fp32 (the reference's precision), 56 031 steps."""
    import somax_b200 as sb
    n, H0 = 32, 500.0
    m = sb.NonlinearShallowWater2D.create(nx=n, ny=n, Lx=1e6, Ly=1e6, g=9.81, f0=1e-4, beta=1.6e-11, H0=H0,
                                          lateral_viscosity=3000.0, bottom_drag=1e-5, wind_amplitude=1e-5,
                                          wind_profile="doublegyre", bc="wall", dtype="float32")
    h0 = np.full((n + 2, n + 2), H0, np.float32)
    dt = 0.2 * m.grid.dx / np.sqrt(9.81 * H0)
    sol = m.integrate(sb.NonlinearSW2DState(h=h0, u=np.zeros_like(h0), v=np.zeros_like(h0)), 0.0, 5e6, dt,
                      max_steps=500_000)
    h, u, v = sol.ys.h[0], sol.ys.u[0], sol.ys.v[0]
    assert np.isfinite(h).all()
    assert float(np.abs(u[2:-2, 2:-2]).max()) == pytest.approx(REF["swm17_max_abs_u"], abs=3e-4)
    assert float(np.abs(v[2:-2, 2:-2]).max()) == pytest.approx(REF["swm17_max_abs_v"], abs=3e-4)
    np.testing.assert_allclose(u[-3, -5:-2], REF["swm17_tail_u"], atol=1e-3)
    np.testing.assert_allclose(v[-3, -5:-2], REF["swm17_tail_v"], atol=1e-3)
    np.testing.assert_allclose(h[-3, -5:-2] - H0, REF["swm17_tail_eta"], atol=1e-3)
    d = m.diagnose(sb.NonlinearSW2DState(h=h, u=u, v=v))
    area = m.grid.dx * m.grid.dy
    # the notebook's "Final KE" is diag.energy of the somax version it was run with (no cell area)
    assert float(d.energy) / area == pytest.approx(REF["swm17_energy_printed"], rel=5e-3)


def test_swm_loss_gradients_match_reference_printout():
    """step17 cell 18: d sum(u^2) / d(viscosity, wind_amplitude) after t1 = 1e5 s (1121 steps),
    printed by eqx.filter_grad; here central differences of the fp64 CUDA path."""
    import somax_b200 as sb
    n, H0 = 32, 500.0

    def loss(nu, tau):
        m = sb.NonlinearShallowWater2D.create(nx=n, ny=n, Lx=1e6, Ly=1e6, g=9.81, f0=1e-4, beta=1.6e-11, H0=H0,
                                              lateral_viscosity=nu, bottom_drag=1e-5, wind_amplitude=tau,
                                              wind_profile="doublegyre", bc="wall", dtype="float64")
        h0 = np.full((n + 2, n + 2), H0)
        dt = 0.2 * m.grid.dx / np.sqrt(9.81 * H0)
        u = m.integrate(sb.NonlinearSW2DState(h=h0, u=np.zeros_like(h0), v=np.zeros_like(h0)), 0.0, 1e5, dt,
                        max_steps=10_000).ys.u[0]
        return float(np.sum(u ** 2))

    g_nu = (loss(3001.0, 1e-5) - loss(2999.0, 1e-5)) / 2.0
    g_tau = (loss(3000.0, 1e-5 + 1e-8) - loss(3000.0, 1e-5 - 1e-8)) / 2e-8
    assert g_tau == pytest.approx(REF["swm17_dloss_dwind"], rel=1e-5)
    assert g_nu == pytest.approx(REF["swm17_dloss_dviscosity"], rel=3e-4)
