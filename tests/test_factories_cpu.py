"""CPU: model construction and test-case factories of the host mirror (hot-path table rows a18,
a19) against the oracle's restatement of the same reference code
(qg/baroclinic.py:277-332, qg/barotropic.py:216-248, swm/multilayer.py:313-377,
gfd_testcases.py:184-362) and the reference's own construction tests
(tests/models/test_gfd_testcases.py, tests/models/test_qg_baroclinic.py:40-76).  No kernel runs:
the device handles are created lazily, on the first compute call."""
import numpy as np
import pytest

import somax_b200 as sb
from oracle import qg as oqg
from oracle import swm as oswm
from oracle import testcases as ot
from somax_b200 import gfd_testcases as g


def test_baroclinic_qg_create_matches_oracle_inputs():
    kw = dict(nx=48, ny=40, lateral_viscosity=15.0, bottom_drag=1e-7, wind_amplitude=1.3e-10)
    m, o = sb.BaroclinicQG.create(**kw), oqg.create_baroclinic(**kw)
    assert m.grid.Nx == 50 and m.grid.Ny == 42 and m.grid.dx == o.dx and m.grid.dy == o.dy
    for a, b in ((m.modal.Cl2m, o.Cl2m), (m.modal.Cm2l, o.Cm2l), (m.helmholtz_lambdas, o.lambdas),
                 (m.beta_y, o.beta_y), (m.wind_forcing, o.wind)):
        assert np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64))
    assert m._H0 == o.H0 == 400.0 and m.consts.n_layers == 3 and m.poisson_bc == "dst"
    assert np.all(np.diff(m.modal.eigenvalues) > 0)                    # ascending (tests/core/test_transforms.py:139-142)
    assert np.allclose(m.modal.Cm2l @ m.modal.Cl2m, np.eye(3), atol=1e-12)
    # wind: -sin(2 pi y / Ly) on rows j*dy; beta_y antisymmetric about the mid-basin row
    assert np.allclose(m.wind_forcing[:, 0], -np.sin(2 * np.pi * np.arange(42) * m.grid.dy / 4e6))
    assert np.all(m.beta_y[20] == 0) and np.allclose(m.beta_y[20 + 5], -m.beta_y[20 - 5])   # beta (y - Ly/2), y = j dy


def test_create_rejects_inconsistent_layers_and_unsupported_options():
    with pytest.raises(ValueError, match="must all be equal"):
        sb.BaroclinicQG.create(n_layers=3, H=(1.0, 2.0), g_prime=(9.81, 0.02, 0.01))
    with pytest.raises(ValueError, match="same length"):
        sb.StratificationProfile.from_layers(H=[1.0, 2.0], g_prime=[9.81])
    with pytest.raises(NotImplementedError, match="dst"):
        sb.BaroclinicQG.create(poisson_bc="dirichlet")


def test_barotropic_is_the_single_layer_case():
    m, o = sb.BarotropicQG.create(nx=32, ny=32, wind_amplitude=1e-12), oqg.create_barotropic(nx=32, ny=32, wind_amplitude=1e-12)
    assert m._H0 == 1.0 and m._engine.nl == 1 and np.array_equal(m._engine.lambdas, [0.0])
    assert np.array_equal(m.beta_y, o.beta_y) and np.array_equal(m.wind_forcing, o.wind)


def test_multilayer_swm_create_matches_oracle_inputs():
    kw = dict(nx=24, ny=20, Lx=1e6, Ly=1e6, f0=1e-4, beta=1.6e-11, n_layers=2, H=(500.0, 4500.0),
              g_prime=(9.81, 0.025), lateral_viscosity=100.0, bottom_drag=1e-7)
    m, o = sb.MultilayerShallowWater2D.create(**kw), oswm.create_multilayer(**kw)
    assert np.array_equal(m.f_field, o.f_field) and np.array_equal(np.asarray(m.strat.g_prime), o.g_prime)
    assert np.array_equal(m.wind_stress_x, o.wind_x) and np.array_equal(m.wind_stress_y, o.wind_y)
    assert m.grid.dx == o.dx and m.bc_type == "periodic"


def test_factories_match_oracle_initial_states():
    m, s = g.doublegyre_baroclinic_qg(nx=32, ny=32)
    assert isinstance(m, sb.BaroclinicQG) and s.q.shape == (3, 34, 34) and s.q.dtype == np.float32 and not s.q.any()
    m, s = g.doublegyre_qg(nx=32, ny=32)
    assert isinstance(m, sb.BarotropicQG) and s.q.shape == (34, 34) and not s.q.any()
    for dtype in ("float32", "float64"):
        m, s = g.baroclinic_instability_swm(nx=32, ny=32, dtype=dtype)
        _, (h, u, v) = ot.baroclinic_instability_swm(nx=32, ny=32, dtype=np.dtype(dtype).type)
        assert np.array_equal(s.h, h) and np.array_equal(s.u, u) and np.array_equal(s.v, v)
        assert s.h.dtype == np.dtype(dtype) and m.dtype == np.dtype(dtype)
    # jet: +0.5 / -0.5 m/s at the basin centre, layer thicknesses H_k, small v perturbation
    assert np.isclose(s.u[0].max(), 0.5, atol=1e-3) and np.isclose(s.u[1].min(), -0.5, atol=1e-3)
    assert np.all(s.h[0] == 500.0) and np.all(s.h[1] == 4500.0) and 0 < np.abs(s.v).max() <= 0.01
    m, s = g.barotropic_jet_instability(nx=32, ny=32)
    assert isinstance(m, sb.NonlinearShallowWater2D) and s.h.shape == (34, 34) and np.all(s.h == 1000.0)
    assert np.array_equal(g.synthetic_qg_state(3, 16, 12), ot.synthetic_qg_state(3, 16, 12))


def test_registry_adapters_build_the_same_models():
    from somax_b200.cli import get_adapter
    blocks = dict(grid={"nx": 32, "ny": 32, "Lx": 4e6, "Ly": 4e6}, consts={"f0": 9.375e-5, "beta": 1.754e-11, "n_layers": 3},
                  stratification={"H": [400.0, 1100.0, 2600.0], "g_prime": [9.81, 0.025, 0.0125]},
                  params={"lateral_viscosity": 15.0, "bottom_drag": 1e-7, "wind_amplitude": 1.3e-10})
    m, s = get_adapter("doublegyre_baroclinic_qg")(**blocks)
    d, _ = g.doublegyre_baroclinic_qg(nx=32, ny=32)
    assert np.array_equal(m.helmholtz_lambdas, d.helmholtz_lambdas) and np.array_equal(m.beta_y, d.beta_y)
    assert m.params.lateral_viscosity == 15.0 and s.q.shape == (3, 34, 34)
    bad = dict(blocks, params={"lateral_viscosity": 15.0})
    with pytest.raises(KeyError, match="bottom_drag"):
        get_adapter("doublegyre_baroclinic_qg")(**bad)
    m64, _ = get_adapter("doublegyre_baroclinic_qg")(**dict(blocks, grid=dict(blocks["grid"], dtype="float64")))
    assert m64.dtype == np.float64


def test_step_plan_and_integrate_argument_checks():
    from somax_b200.core import SaveAt, step_plan
    assert step_plan(0.0, 3000.0, 600.0) == (5, 0.0)
    n, rem = step_plan(0.0, 3100.0, 600.0)
    assert n == 5 and np.isclose(rem, 100.0)
    with pytest.raises(ValueError):
        step_plan(1.0, 0.0, 1.0)
    m, s = g.doublegyre_qg(nx=16, ny=16)
    with pytest.raises(NotImplementedError, match="Tsit5"):
        m.integrate(s, 0.0, 10.0, 1.0, solver=type("Dopri5", (), {})())
    with pytest.raises(RuntimeError, match="max_steps"):
        m.integrate(s, 0.0, 1e6, 1.0)                                  # 1e6 steps > the 4096 default, checked before any launch
    with pytest.raises(TypeError, match="unsupported integrate"):
        m.integrate(s, 0.0, 10.0, 1.0, adjoint=None)
    assert SaveAt(t1=True).ts is None
