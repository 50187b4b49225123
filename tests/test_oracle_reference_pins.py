"""The oracle against numbers the REFERENCE printed (tests/golden/reference_notebook_outputs.json,
extracted from the executed tutorials under /root/reference/content/tutorials by
tools/extract_notebook_outputs.py).  These are the only reference-produced values for the
path; they pin

  * the elliptic solve behind BaroclinicQG._invert_pv / BarotropicQG._invert_pv
    (PoissonSolver2D / HelmholtzSolver2D call the same finitevolx function,
    somax/_src/models/pde2d/poisson.py:36-38,129-133): DST-I, finite-difference eigenvalues,
    every point of the array an unknown;
  * the shallow-water operator set (NonlinearShallowWater2D.vector_field /
    apply_boundary_conditions, somax/_src/models/swm/nonlinear_2d.py:132-208) through 1121 Tsit5
    steps: upwind advection written on [2:-2, 2:-2], flux-form diffusion with zero ghost fluxes,
    wall BCs, wind, Coriolis / PV fluxes, Bernoulli gradient, Tsit5 itself;

and they show that the other settings of ``OperatorSpec`` do NOT reproduce the reference.
"""
import json
import os

import numpy as np
import pytest

from oracle import swm as oswm
from oracle.elliptic import helmholtz_dst
from oracle.operators import OperatorSpec

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_notebook_outputs.json")))["parsed"]


def _poisson_case(n, spec, mode=(1, 1)):
    """content/tutorials/step09_laplace_2d.py:112-122,170-176: x = arange(Nx) * dx, dx = 1 / n."""
    dx = 1.0 / n
    x = np.arange(n + 2) * dx
    X, Y = np.meshgrid(x, x)
    a, b = mode
    exact = np.sin(a * np.pi * X) * np.sin(b * np.pi * Y)
    rhs = -(a * a + b * b) * np.pi ** 2 * exact
    phi = helmholtz_dst(rhs, dx, dx, 0.0, spec)
    err = phi[1:-1, 1:-1] - exact[1:-1, 1:-1]
    return float(np.sqrt(np.mean(err ** 2))), float(np.abs(err).max())


def test_poisson_errors_match_reference_printout():
    spec = OperatorSpec()
    for n, ref in REF["poisson_dirichlet_l2_by_n"].items():
        l2, _ = _poisson_case(int(n), spec)
        assert l2 == pytest.approx(ref, rel=1e-5), (n, l2, ref)   # reference ran in fp32
    l2, linf = _poisson_case(64, spec)
    assert l2 == pytest.approx(REF["poisson_dirichlet_l2_n64"], rel=1e-5)
    assert linf == pytest.approx(REF["poisson_dirichlet_linf_n64"], rel=1e-5)
    l2, _ = _poisson_case(64, spec, mode=(2, 3))
    assert l2 == pytest.approx(REF["poisson_dirichlet_mode23_l2_n64"], rel=1e-5)


@pytest.mark.parametrize("kw", [dict(dst_fd_eigenvalues=False), dict(dst_full_array=False),
                                dict(dst_fd_eigenvalues=False, dst_full_array=False)])
def test_other_elliptic_conventions_do_not_match(kw):
    l2, linf = _poisson_case(64, OperatorSpec(**kw))
    assert abs(l2 / REF["poisson_dirichlet_l2_n64"] - 1.0) > 1e-3
    l2, _ = _poisson_case(64, OperatorSpec(**kw), mode=(2, 3))
    assert abs(l2 / REF["poisson_dirichlet_mode23_l2_n64"] - 1.0) > 1e-3


def test_helmholtz_maxima_match_reference_printout():
    """content/tutorials/step11_helmholtz_2d.py:70-96 (printed with 6 decimals)."""
    n = 64
    dx = 1.0 / n
    x = np.arange(n + 2) * dx
    X, Y = np.meshgrid(x, x)
    rhs = np.sin(np.pi * X) * np.sin(np.pi * Y)
    for lam, ref in REF["helmholtz_maxabs_by_lambda"].items():
        phi = helmholtz_dst(rhs, dx, dx, float(lam))
        assert float(np.abs(phi[1:-1, 1:-1]).max()) == pytest.approx(ref, abs=6e-7), lam
    phi = helmholtz_dst(rhs, dx, dx, 0.0, OperatorSpec(dst_full_array=False))
    assert abs(float(np.abs(phi[1:-1, 1:-1]).max()) - REF["helmholtz_maxabs_by_lambda"]["0.0"]) > 1e-3


def _swm17(spec, nu=3000.0, tau=1e-5, t1=1e5, dtype=np.float64):
    """content/tutorials/step17_shallow_water_2d.ipynb cells 3-7, 18."""
    n, H0 = 32, 500.0
    m = oswm.create_multilayer(nx=n, ny=n, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, n_layers=1, H=(1.0,),
                               g_prime=(9.81,), lateral_viscosity=nu, bottom_drag=1e-5,
                               wind_amplitude=tau, wind_profile="doublegyre", bc="wall", spec=spec)
    h0 = np.full((1, n + 2, n + 2), H0, dtype)
    dt = 0.2 * m.dx / np.sqrt(9.81 * H0)
    return m, m.integrate(h0, np.zeros_like(h0), np.zeros_like(h0), 0.0, t1, dt)


def _loss(spec, nu, tau):
    _, (_, u, _) = _swm17(spec, nu, tau)
    return float(np.sum(u ** 2))


def test_swm_gradients_match_reference_printout():
    """d sum(u^2) / d(viscosity, wind_amplitude) after 1121 steps, printed by eqx.filter_grad in
    fp32; central differences of the fp64 oracle."""
    spec = OperatorSpec()
    g_nu = (_loss(spec, 3001.0, 1e-5) - _loss(spec, 2999.0, 1e-5)) / 2.0
    g_tau = (_loss(spec, 3000.0, 1e-5 + 1e-8) - _loss(spec, 3000.0, 1e-5 - 1e-8)) / 2e-8
    assert g_tau == pytest.approx(REF["swm17_dloss_dwind"], rel=1e-5)
    assert g_nu == pytest.approx(REF["swm17_dloss_dviscosity"], rel=3e-4)


def test_swm_other_conventions_do_not_match():
    lap = OperatorSpec(diffusion_flux_form=False)
    g_nu = (_loss(lap, 3001.0, 1e-5) - _loss(lap, 2999.0, 1e-5)) / 2.0
    assert abs(g_nu / REF["swm17_dloss_dviscosity"] - 1.0) > 5e-3
    reg1 = OperatorSpec(advection_region2=False)
    g_tau = (_loss(reg1, 3000.0, 1e-5 + 1e-8) - _loss(reg1, 3000.0, 1e-5 - 1e-8)) / 2e-8
    assert abs(g_tau / REF["swm17_dloss_dwind"] - 1.0) > 1e-2


@pytest.mark.skipif(not os.environ.get("SOMAX_B200_SLOW"), reason="56 031 oracle steps (4 min); the same run is a -m gpu test")
def test_swm_spinup_matches_reference_printout():
    m, (h, u, v) = _swm17(OperatorSpec(), t1=5e6)
    assert float(np.abs(u[0, 2:-2, 2:-2]).max()) == pytest.approx(REF["swm17_max_abs_u"], abs=2e-4)
    assert float(np.abs(v[0, 2:-2, 2:-2]).max()) == pytest.approx(REF["swm17_max_abs_v"], abs=2e-4)
    np.testing.assert_allclose(u[0, -3, -5:-2], REF["swm17_tail_u"], atol=5e-4)
    np.testing.assert_allclose(v[0, -3, -5:-2], REF["swm17_tail_v"], atol=5e-4)
    np.testing.assert_allclose(h[0, -3, -5:-2] - 500.0, REF["swm17_tail_eta"], atol=5e-4)
