"""CPU: the oracle restatement against the reference's own property / known-answer tests.

Each test names the reference test it ports.  These are the only pins the reference offers for
this path: parity is otherwise unpinned (no golden vectors exist).
"""
import numpy as np
import pytest

from oracle import operators as op
from oracle import qg, swm, testcases, tsit5
from oracle.elliptic import helmholtz_dst
from oracle.modal import ModalTransform


# ---- tests/models/test_poisson.py -------------------------------------------------------
def test_poisson_sinusoidal_dirichlet():  # test_poisson.py:11-27
    nx = ny = 64
    dx = dy = 1.0 / nx
    x = np.arange(nx + 2) * dx
    X, Y = np.meshgrid(x, x)
    rhs = -2.0 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y)
    phi = helmholtz_dst(rhs[None], dx, dy, np.zeros(1))[0]
    exact = np.sin(np.pi * X) * np.sin(np.pi * Y)
    err = np.sqrt(np.mean((phi[1:-1, 1:-1] - exact[1:-1, 1:-1]) ** 2))
    assert err < 0.1


def test_poisson_zero_rhs():  # test_poisson.py:29-34
    phi = helmholtz_dst(np.zeros((1, 34, 34)), 1.0, 1.0, np.zeros(1))
    assert np.abs(phi).max() < 1e-10


def test_helmholtz_lambda0_is_poisson_and_amplitude_decreases():  # test_poisson.py:67-90
    rng = np.random.default_rng(0)
    rhs = np.zeros((1, 34, 34))
    rhs[0, 1:-1, 1:-1] = rng.standard_normal((32, 32))
    a = helmholtz_dst(rhs, 0.1, 0.1, np.zeros(1))
    b = helmholtz_dst(rhs, 0.1, 0.1, 0.0)
    assert np.abs(a - b).max() < 1e-10
    c = helmholtz_dst(rhs, 0.1, 0.1, np.array([50.0]))
    assert np.abs(c).max() < np.abs(a).max()


def test_helmholtz_inverts_five_point_operator():  # SURVEY App. B (checked numerically there)
    rng = np.random.default_rng(1)
    rhs = np.zeros((2, 20, 26))
    rhs[:, 1:-1, 1:-1] = rng.standard_normal((2, 18, 24))
    lam = np.array([0.3, 2.0])
    psi = helmholtz_dst(rhs, 0.7, 0.9, lam)
    res = op.laplacian(psi, 0.7, 0.9) - lam[:, None, None] * psi
    assert np.abs(res[:, 1:-1, 1:-1] - rhs[:, 1:-1, 1:-1]).max() < 1e-12


# ---- tests/core/test_transforms.py -------------------------------------------------------
def test_modal_roundtrip_and_inverse():  # test_transforms.py:119-137
    t = ModalTransform.from_physics((500.0, 4500.0), (9.81, 0.025), 1e-4)
    x = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert np.allclose(t.to_layer(t.to_modal(x)), x, atol=1e-5)
    assert np.allclose(t.Cl2m @ t.Cm2l, np.eye(2), atol=1e-5)
    assert np.all(t.rossby_radii > 0)


def test_eigenvalues_sorted():  # test_transforms.py:139-142
    t = ModalTransform.from_physics((400.0, 1100.0, 2600.0), (9.81, 0.025, 0.0125), 9.375e-5)
    assert np.all(t.eigenvalues[:-1] <= t.eigenvalues[1:])


# ---- tests/models/test_qg_baroclinic.py ---------------------------------------------------
def _bcqg(**kw):
    return qg.create_baroclinic(nx=16, ny=16, **kw)


def test_qg_zero_pv_zero_tendency():  # test_qg_baroclinic.py:79-94
    m = _bcqg()
    dq = m.rhs(np.zeros((3, 18, 18)))
    assert np.abs(dq).max() < 1e-12


def test_qg_wind_top_layer_only():  # test_qg_baroclinic.py:96-112
    m = _bcqg(wind_amplitude=1e-10)
    dq = m.rhs(np.zeros((3, 18, 18)))
    assert np.abs(dq[0]).max() > 0
    assert np.abs(dq[1:]).max() < 1e-20


def test_qg_bc_zeroes_ring_only():  # test_qg_baroclinic.py:166-183
    m = _bcqg()
    q = np.random.default_rng(0).standard_normal((3, 18, 18))
    b = m.bc(q)
    assert np.all(b[:, 0] == 0) and np.all(b[:, -1] == 0)
    assert np.all(b[:, :, 0] == 0) and np.all(b[:, :, -1] == 0)
    assert np.array_equal(b[:, 1:-1, 1:-1], q[:, 1:-1, 1:-1])


def test_qg_wind_spins_up_pv():  # test_qg_baroclinic.py:202-223
    m = _bcqg(lateral_viscosity=100.0, wind_amplitude=1e-5)
    q = m.integrate(np.zeros((3, 18, 18), np.float32), 0.0, 100.0, 1.0)
    assert np.isfinite(q).all()
    assert np.abs(q[0, 2:-2, 2:-2]).max() > 1e-10


def test_barotropic_zero_and_finite():  # test_qg_barotropic.py:40-60,86-95
    m = qg.create_barotropic(nx=16, ny=16, lateral_viscosity=100.0, wind_amplitude=1e-5)
    assert np.abs(qg.create_barotropic(nx=16, ny=16).rhs(np.zeros((1, 18, 18)))).max() < 1e-12
    q = m.integrate(np.zeros((1, 18, 18), np.float32), 0.0, 100.0, 1.0)
    assert np.isfinite(q).all() and np.abs(q[0, 2:-2, 2:-2]).max() > 1e-10


def test_ring_drifts_with_wind_only():
    m = _bcqg(wind_amplitude=1e-9)
    q = m.integrate(np.zeros((3, 18, 18)), 0.0, 10.0, 1.0)
    expect = 10.0 * m.tau0 * m.wind[0, 0] / m.H0
    assert q[0, 0, 0] == pytest.approx(expect, rel=1e-12, abs=1e-30)
    assert np.all(q[1:, 0, :] == 0)


# ---- tests/models/test_swm_multilayer.py ---------------------------------------------------
def _swm(**kw):
    return swm.create_multilayer(nx=16, ny=16, Lx=1e6, Ly=1e6, n_layers=2, H=(500.0, 4500.0),
                                 g_prime=(9.81, 0.025), **kw)


def _rest(m):
    h = np.ones((2, 18, 18)) * m.H[:, None, None]
    return h, np.zeros_like(h), np.zeros_like(h)


def test_swm_rest_zero_tendency():  # test_swm_multilayer.py:94-102
    m = _swm()
    t = m.rhs(*m.bc(*_rest(m)))
    assert max(np.abs(a).max() for a in t) < 1e-10


def test_swm_wind_top_layer_only():  # test_swm_multilayer.py:104-113
    m = _swm(wind_amplitude=1e-5)
    _, du, _ = m.rhs(*m.bc(*_rest(m)))
    assert np.abs(du[0]).max() > 0 and np.abs(du[-1]).max() < 1e-15


def test_swm_pressure_coupling_reaches_layer1():  # test_swm_multilayer.py:130-141
    m = _swm()
    h, u, v = _rest(m)
    h[0, 8:10, 8:10] += 1.0
    _, du, dv = m.rhs(*m.bc(h, u, v))
    assert np.abs(du[1]).max() > 0 or np.abs(dv[1]).max() > 0


def test_swm_wall_bc_u_zero_at_x_edges():  # test_swm_multilayer.py:163-172
    m = _swm(bc="wall")
    rng = np.random.default_rng(0)
    h, u, v = (rng.standard_normal((2, 18, 18)) for _ in range(3))
    hb, ub, vb = m.bc(h, u, v)
    assert np.all(ub[:, :, 0] == 0) and np.all(ub[:, :, -1] == 0) and np.all(ub[:, :, -2] == 0)
    assert np.all(vb[:, 0, :] == 0) and np.all(vb[:, -1, :] == 0) and np.all(vb[:, -2, :] == 0)
    assert np.array_equal(hb[:, 0, 1:-1], hb[:, 1, 1:-1])


def test_swm_periodic_ring():  # tests/models/test_navier_stokes.py:72-78 (enforce_periodic)
    x = np.random.default_rng(0).standard_normal((18, 18))
    p = op.enforce_periodic(x)
    assert np.array_equal(p[:, 0], p[:, -2]) and np.array_equal(p[0, :], p[-2, :])


# ---- tests/test_cli_run.py (behavioural pins) -----------------------------------------------
def test_swm_jet_64_dt300_diverges_but_32_dt10_is_finite():  # test_cli_run.py:158-212
    m, (h, u, v) = testcases.baroclinic_instability_swm(nx=64, ny=64)
    with np.errstate(all="ignore"):
        out = m.integrate(h, u, v, 0.0, 48 * 300.0, 300.0)
    assert not all(np.isfinite(a).all() for a in out)
    m, (h, u, v) = testcases.baroclinic_instability_swm(nx=32, ny=32)
    out = m.integrate(h, u, v, 0.0, 3600.0, 10.0)
    assert all(np.isfinite(a).all() for a in out)


# ---- Arakawa Jacobian (SURVEY App. B.3 checks) ----------------------------------------------
def test_arakawa_antisymmetric_and_converges():
    n = 64
    d = 1.0 / n
    x = np.arange(n + 2) * d
    X, Y = np.meshgrid(x, x)
    f = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    g = np.cos(4 * np.pi * X) * np.sin(2 * np.pi * Y)
    J = op.arakawa_jacobian(f, g, d, d)
    assert np.abs(J + op.arakawa_jacobian(g, f, d, d)).max() < 1e-12
    fx = 2 * np.pi * np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    fy = -2 * np.pi * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y)
    gx = -4 * np.pi * np.sin(4 * np.pi * X) * np.sin(2 * np.pi * Y)
    gy = 2 * np.pi * np.cos(4 * np.pi * X) * np.cos(2 * np.pi * Y)
    exact = (fx * gy - fy * gx)[1:-1, 1:-1]
    assert np.abs(J - exact).max() < 0.05 * np.abs(exact).max()


# ---- Tsit5 -------------------------------------------------------------------------------
def test_tsit5_tableau_and_exponential_decay():  # tests/core/test_model.py:12-74 (toy model)
    for r, c in zip(tsit5.A[1:], tsit5.C[1:]):
        assert sum(r) == pytest.approx(c, abs=1e-14)
    y = tsit5.integrate(lambda y: (-y[0],), lambda y: y, (np.array([1.0]),), 0.0, 1.0, 0.01)[0]
    assert y[0] == pytest.approx(np.exp(-1.0), rel=1e-10)
    assert tsit5.step_plan(0.0, 1.0, 0.3) == (3, pytest.approx(0.1))
    assert tsit5.step_plan(0.0, 86400.0 * 30, 600.0) == (4320, 0.0)


# ---------------------------------------------------------------- reparameterized QG
def _reparam(nl=2, **kw):
    from oracle.reparam import create_reparameterized
    args = dict(nx=16, ny=16, n_layers=2, H=(500.0, 500.0), g_prime=(9.81, 0.02))
    if nl == 3:
        args = dict(nx=16, ny=16)
    args.update(kw)
    m = create_reparameterized(**args)
    h = np.ones((m.swm.nl, 18, 18)) * m.swm.H[:, None, None]
    return m, h, np.zeros_like(h), np.zeros_like(h)


def test_reparameterized_qg_reference_properties():
    """tests/models/test_qg_reparameterized.py:60-135: rest state is a fixed point of the projection
    and of the flow; the projection is (nearly) idempotent and produces geostrophic velocities;
    curl(grad_perp(psi)) is the 5-point Laplacian, which is what makes (Q G)^-1 the Helmholtz solve."""
    from oracle import operators as op
    from oracle.reparam import grad_perp
    m, h, u, v = _reparam()
    ph, pu, pv = m.project(h, u, v)
    assert np.abs(ph - h).max() < 1e-6 and np.abs(pu).max() < 1e-10 and np.abs(pv).max() < 1e-10
    assert all(np.abs(a).max() < 1e-10 for a in m.rhs(*m.bc(h, u, v)))
    h2 = h.copy()
    h2[0, 9, 9] += 1.0
    p1 = m.project(h2, u, v)
    p2 = m.project(*p1)
    assert all(np.abs(a - b).max() < 0.1 for a, b in zip(p1, p2))
    assert np.abs(p1[1]).max() > 0 and np.abs(p1[2]).max() > 0 and all(np.isfinite(a).all() for a in p1)
    psi = np.random.default_rng(0).standard_normal((2, 18, 18))
    ug, vg = grad_perp(psi, m.swm.dx, m.swm.dy)
    lap = op.laplacian(psi, m.swm.dx, m.swm.dy)
    s = (slice(None), slice(2, -2), slice(2, -2))       # away from the zero ring of u_g, v_g
    assert np.allclose(op.curl(ug, vg, m.swm.dx, m.swm.dy)[s], lap[s], rtol=1e-10, atol=1e-22)
    out = m.integrate(h2, u, v, 0.0, 600.0, 60.0)
    assert all(np.isfinite(a).all() for a in out)
    m3, h3, u3, v3 = _reparam(nl=3, wind_amplitude=8e-5, lateral_viscosity=15.0)
    out = m3.integrate(h3, u3, v3, 0.0, 600.0, 60.0)
    assert all(np.isfinite(a).all() for a in out) and np.abs(out[1][0]).max() > 0
    with pytest.raises(ValueError, match="wall"):
        _reparam(bc="periodic")
