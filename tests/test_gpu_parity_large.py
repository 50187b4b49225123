"""Oracle parity of stepping and of the raw vector field at sizes where the CUDA path runs its
production kernels: CTAs of the TMA stencil that do not touch the ring (predicate-free path), the
radix-8 row FFT at 512 / 1024 and the three-pass register FFT at nx >= 4096, fp64-carry (low-k) and
plain Thomas strips under stepping - and at BASELINE.json's configurations C2 (3 x 128^2, 100
steps), C5 (256^2 members of an ensemble) and C3 (shallow water at 256^2).  Tolerances are
BASELINE.json's: relative L2 <= 1e-5 in fp32, <= 1e-12 in fp64; fp32 results are compared with the
fp64 oracle.
"""
import numpy as np
import pytest

from test_gpu_parity import TOL, assert_swm_parity, qg_pair, qstate, rel, swm_pair, swm_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nx,ny", [(512, 512), (1024, 1024), (4096, 128)])
def test_qg_vector_field_and_steps_at_size(nx, ny, dtype):
    import somax_b200 as sb
    om, gm = qg_pair(nx, ny, dtype)
    om.workers = -1
    q0 = qstate(3, nx, ny, dtype, ring=True)
    st = sb.BaroclinicQGState(q=q0)
    q64 = q0.astype(np.float64)
    # raw vector field and _rhs = vector_field(BC(q)), per layer (the lower layers are not
    # hidden behind the wind-forced top layer)
    # fp32: differences of psi over one cell lose ~ nx * eps in ANY fp32 evaluation of the Jacobian
    # (the reference's included); the oracle run in fp32 measures that floor (as for the
    # shallow-water momentum tendencies, test_gpu_parity.assert_swm_parity)
    for got, ref, same in ((gm.vector_field(0.0, st).q, om.rhs(q64), om.rhs(q0)),
                           (gm.build_terms().vf(0.0, st).q, om.rhs(om.bc(q64)), om.rhs(om.bc(q0)))):
        for l in range(3):
            tol = max(2e-5, 1.5 * rel(same[l], ref[l])) if dtype == np.float32 else 1e-11
            assert rel(got[l], ref[l]) <= tol, l
    dt = 600.0 * 128 / nx
    for steps in (1, 10):
        got = gm.integrate(st, 0.0, steps * dt, dt).ys.q[0]
        ref = om.integrate(q64, 0.0, steps * dt, dt)
        assert rel(got, ref) <= TOL[dtype], steps
        for l in range(3):
            assert rel(got[l], ref[l]) <= 2 * TOL[dtype], (steps, l)
        assert rel(gm._invert_pv(got), om.invert_pv(ref)) <= 2 * TOL[dtype]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_qg_c2_128_100_steps(dtype):
    """BASELINE config 2: doublegyre_bc_qg (3 x 128^2, dt = 600 s), 1 and 100 steps, diagnostics."""
    import somax_b200 as sb
    om, gm = qg_pair(128, 128, dtype)
    q0 = qstate(3, 128, 128, dtype, ring=True)
    for steps in (1, 100):
        got = gm.integrate(sb.BaroclinicQGState(q=q0), 0.0, steps * 600.0, 600.0).ys.q[0]
        ref = om.integrate(q0.astype(np.float64), 0.0, steps * 600.0, 600.0)
        assert rel(got, ref) <= TOL[dtype], steps
        assert rel(gm._invert_pv(got), om.invert_pv(ref)) <= 2 * TOL[dtype]
    d, dref = gm.diagnose(sb.BaroclinicQGState(q=got)), om.diagnose(ref)
    assert np.allclose(d.kinetic_energy, dref["kinetic_energy"], rtol=1e-4)
    assert np.allclose(d.enstrophy, dref["enstrophy"], rtol=1e-4)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_qg_c5_ensemble_of_256_members(dtype):
    """BASELINE config 5's member shape: 4 members of 3 x 256^2 stepped as one batch, each
    against the oracle (1 and 10 steps)."""
    import somax_b200 as sb
    from oracle.testcases import synthetic_qg_state
    om, gm = qg_pair(256, 256, dtype)
    qs = np.stack([synthetic_qg_state(3, 256, 256, seed=10_000 + e, dtype=np.float64) for e in range(4)]).astype(dtype)
    dt = 300.0
    for steps in (1, 10):
        got = gm.integrate(sb.BaroclinicQGState(q=qs), 0.0, steps * dt, dt).ys.q[0]
        assert got.shape == qs.shape
        for e in range(4):
            ref = om.integrate(qs[e].astype(np.float64), 0.0, steps * dt, dt)
            assert rel(got[e], ref) <= TOL[dtype], (steps, e)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("bc", ["periodic", "wall"])
def test_swm_c3_256_10_steps(bc, dtype):
    """BASELINE config 3 (swm_jet) at 256^2: 1 and 10 steps of the factory state."""
    import somax_b200 as sb
    om, gm = swm_pair(256, 256, dtype, bc)
    h, u, v = swm_state(256, 256, dtype, noise=False)
    dt = 20.0 * 64 / 256
    for steps in (1, 10):
        sol = gm.integrate(sb.MultilayerSW2DState(h=h, u=u, v=v), 0.0, steps * dt, dt)
        ref = om.integrate(*[a.astype(np.float64) for a in (h, u, v)], 0.0, steps * dt, dt)
        same = om.integrate(h, u, v, 0.0, steps * dt, dt)
        for name, r, rs in zip("huv", ref, same):
            assert_swm_parity(getattr(sol.ys, name)[0], r, rs, dtype, name)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_barotropic_keeps_the_ring_of_psi(dtype):
    """BarotropicQG._invert_pv (qg/barotropic.py:113-121) returns the solver output with its
    ring; BaroclinicQG zeroes it.  Both solve on the whole array."""
    from oracle import qg as oqg
    import somax_b200 as sb
    rng = np.random.default_rng(2)
    for nx, ny, solver in ((64, 48, 1), (30, 20, 2), (256, 256, 1)):
        kw = dict(nx=nx, ny=ny)
        om = oqg.create_barotropic(**kw)
        gm = sb.BarotropicQG.create(dtype=np.dtype(dtype).name, solver=solver, **kw)
        q = (1e-6 * rng.standard_normal((ny + 2, nx + 2))).astype(dtype)
        psi, ref = gm._invert_pv(q), om.invert_pv(q.astype(np.float64)[None])[0]
        assert rel(psi, ref) <= (5e-6 if dtype == np.float32 else 1e-12)
        ring = np.ones_like(ref, bool); ring[1:-1, 1:-1] = False
        assert np.abs(ref[ring]).max() > 0
        assert rel(psi[ring], ref[ring]) <= (2e-5 if dtype == np.float32 else 1e-11)
        dq = gm.vector_field(0.0, sb.BarotropicQGState(q=q)).q
        assert rel(dq, om.rhs(q.astype(np.float64)[None])[0]) <= (5e-5 if dtype == np.float32 else 1e-10)


def test_unsupported_dst_conventions_are_refused():
    import somax_b200 as sb
    from somax_b200 import _lib
    for bit in (_lib.SPEC_DST_CONTINUOUS, _lib.SPEC_DST_INTERIOR):
        m = sb.BaroclinicQG.create(nx=16, ny=16, spec=_lib.DEFAULT_SPEC | bit)
        with pytest.raises(_lib.SomaxB200Error, match="only the reference's DST convention"):
            m._invert_pv(np.zeros((3, 18, 18), np.float32))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_diagnose_fields_match_oracle(dtype):
    """Every field of the Diagnostics pytrees (qg/baroclinic.py:197-228, qg/barotropic.py:159-183,
    swm/multilayer.py:225-256), not only the scalar sums."""
    from oracle import qg as oqg
    import somax_b200 as sb
    tol = 5e-6 if dtype == np.float32 else 1e-12
    om, gm = qg_pair(64, 48, dtype)
    q = qstate(3, 64, 48, dtype, ring=True)
    d, r = gm.diagnose(sb.BaroclinicQGState(q=q)), om.diagnose(q.astype(np.float64))
    assert rel(d.psi, r["psi"]) <= tol
    # u, v, zeta difference psi over one cell: (n / pi) times the rounding of psi
    for name, key in (("u", "u"), ("v", "v"), ("relative_vorticity", "relative_vorticity")):
        assert rel(getattr(d, name), r[key]) <= tol * (400 if key == "relative_vorticity" else 20), name
    assert np.allclose(d.kinetic_energy, r["kinetic_energy"], rtol=1e-4)
    assert np.allclose(d.total_enstrophy, r["total_enstrophy"], rtol=1e-4)
    assert np.allclose(d.rossby_radii, r["rossby_radii"], rtol=1e-12)
    # barotropic: psi keeps its ring, so u, v, zeta next to the ring see it
    ob = oqg.create_barotropic(nx=64, ny=64)
    gb = sb.BarotropicQG.create(nx=64, ny=64, dtype=np.dtype(dtype).name)
    qb = qstate(1, 64, 64, dtype, ring=True)[0]
    d, r = gb.diagnose(sb.BarotropicQGState(q=qb)), ob.diagnose(qb.astype(np.float64)[None])
    assert rel(d.psi, r["psi"][0]) <= tol
    assert rel(d.u, r["u"][0]) <= 20 * tol and rel(d.v, r["v"][0]) <= 20 * tol
    assert rel(d.relative_vorticity, r["relative_vorticity"][0]) <= 400 * tol
    assert np.allclose(d.kinetic_energy, r["kinetic_energy"][0], rtol=1e-4)
    # shallow water
    osw, gsw = swm_pair(64, 40, dtype, "wall")
    h, u, v = swm_state(64, 40, dtype)
    d = gsw.diagnose(sb.MultilayerSW2DState(h=h, u=u, v=v))
    r = osw.diagnose(*[a.astype(np.float64) for a in (h, u, v)])
    for name in ("potential_vorticity", "relative_vorticity", "kinetic_energy_field"):
        assert rel(getattr(d, name), r[name]) <= (2e-5 if dtype == np.float32 else 1e-12), name
    assert np.allclose(d.energy, r["energy"], rtol=1e-4) and np.allclose(d.enstrophy, r["enstrophy"], rtol=1e-4)


def test_saveat_semantics_follow_diffrax():
    """SaveAt(ts=...) (core/model.py:75-88): BCs on state0 only - stepping continues from the
    un-projected saved states, so the last save equals the plain t1 run bit for bit; host tensors of
    different save times do not alias; SaveAt(t0=True) returns BC(state0); off-grid ts raise."""
    import torch
    import somax_b200 as sb
    om, gm = qg_pair(32, 32, np.float64)
    q0 = qstate(3, 32, 32, np.float64, ring=True)
    st = sb.BaroclinicQGState(q=q0)
    one = gm.integrate(st, 0.0, 3600.0, 600.0).ys.q[0]
    sol = gm.integrate(st, 0.0, 3600.0, 600.0, saveat=sb.SaveAt(ts=[1200.0, 2400.0, 3600.0]))
    assert sol.ys.q.shape[0] == 3 and np.array_equal(sol.ys.q[2], one)
    assert rel(sol.ys.q[0], om.integrate(q0, 0.0, 1200.0, 600.0)) <= 1e-12
    assert rel(sol.ys.q[2], om.integrate(q0, 0.0, 3600.0, 600.0)) <= 1e-12
    # the ring of an intermediate save is NOT projected (wind forcing drifts it)
    assert np.abs(sol.ys.q[0][0, :, 0]).max() > 0
    pin = torch.as_tensor(q0).pin_memory()
    solp = gm.integrate(sb.BaroclinicQGState(q=pin), 0.0, 3600.0, 600.0, saveat=sb.SaveAt(ts=[1200.0, 3600.0]))
    assert np.array_equal(solp.ys.q[0].numpy(), sol.ys.q[0]) and np.array_equal(solp.ys.q[1].numpy(), one)
    s0 = gm.integrate(st, 0.0, 1200.0, 600.0, saveat=sb.SaveAt(t0=True, t1=True))
    assert s0.ys.q.shape[0] == 2 and np.array_equal(s0.ys.q[0], om.bc(q0)) and np.array_equal(s0.ys.q[1], sol.ys.q[0])
    with pytest.raises(ValueError, match="step grid"):
        gm.integrate(st, 0.0, 3600.0, 600.0, saveat=sb.SaveAt(ts=[900.0]))
    # shallow water: same continuation rule
    osw, gsw = swm_pair(32, 32, np.float64, "wall")
    h, u, v = swm_state(32, 32, np.float64)
    sw = sb.MultilayerSW2DState(h=h, u=u, v=v)
    a = gsw.integrate(sw, 0.0, 80.0, 20.0).ys
    b = gsw.integrate(sw, 0.0, 80.0, 20.0, saveat=sb.SaveAt(ts=[40.0, 80.0])).ys
    for f in "huv":
        assert np.array_equal(getattr(b, f)[1], getattr(a, f)[0]), f


def test_two_host_threads_two_handles():
    """include/somax_b200.h: distinct handles may be used concurrently.  Two host threads, each
    with its own models (QG and shallow water) on its own CUDA stream, step at the same time; the
    results equal the serial ones bit for bit (first use of every kernel happens inside the
    threads, so the per-device attribute set-up races if it is not guarded)."""
    import threading
    import torch
    import somax_b200 as sb
    q0 = qstate(3, 256, 192, np.float32, ring=True)
    h, u, v = swm_state(128, 96, np.float32)

    def work(seed, out):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                gm = sb.BaroclinicQG.create(nx=256, ny=192, lateral_viscosity=15.0, bottom_drag=1e-7, wind_amplitude=1.3e-10)
                sw = swm_pair(128, 96, np.float32, "wall")[1]
                res = []
                for _ in range(3):
                    res.append(gm.integrate(sb.BaroclinicQGState(q=q0 * seed), 0.0, 12 * 300.0, 300.0).ys.q[0])
                    s = sw.integrate(sb.MultilayerSW2DState(h=h, u=u * seed, v=v), 0.0, 12 * 10.0, 10.0).ys
                    res.append(np.stack([s.h[0], s.u[0], s.v[0]]))
                out[seed] = res
        except Exception as e:      # surfaced by the main thread
            out[seed] = e

    par = {}
    ts = [threading.Thread(target=work, args=(sd, par)) for sd in (1.0, 0.5)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    ser = {}
    for sd in (1.0, 0.5):
        work(sd, ser)
    for sd in (1.0, 0.5):
        assert not isinstance(par[sd], Exception), par[sd]
        for a, b in zip(par[sd], ser[sd]):
            assert np.array_equal(a, b)


def test_energy_enstrophy_series_over_a_full_chunk():
    """BASELINE.json: energy / enstrophy time series within 1e-4 over a full run.  C2
    (doublegyre_bc_qg, 3 x 128^2, dt = 600 s) over half of a 30-day output chunk of the reference
    run (2160 steps), and the shallow-water jet at 64^2 over 2000 steps, ten samples each, through the
    fused device diagnostics.  The fp64 pipeline must follow the fp64 oracle to 1e-9 the whole way.
    The fp32 pipeline is held to 1e-4, or - where rounding differences have been amplified by the
    flow beyond that in ANY fp32 evaluation - to 3x the distance between the oracle run in fp32 and
    in fp64 (the 3-layer configuration has an indefinite barotropic Helmholtz mode)."""
    import somax_b200 as sb
    q0 = qstate(3, 128, 128, np.float32, ring=True)
    dt, nstep, every = 600.0, 2160, 216
    ts = [every * dt * (i + 1) for i in range(nstep // every)]

    def oracle_series(om, x0):
        out = []
        om.integrate(x0, 0.0, ts[-1], dt, on_step=lambda i, q: out.append(om.diagnose(q)) if i % every == 0 else None)
        return out

    om, g32 = qg_pair(128, 128, np.float32)
    g64 = qg_pair(128, 128, np.float64)[1]
    ref64, ref32 = oracle_series(om, q0.astype(np.float64)), oracle_series(om, q0)
    assert len(ref64) == len(ts)
    for gm, x0, is32 in ((g64, q0.astype(np.float64), False), (g32, q0, True)):
        sol = gm.integrate(sb.BaroclinicQGState(q=x0), 0.0, ts[-1], dt, saveat=sb.SaveAt(ts=ts), max_steps=None)
        for i, dref in enumerate(ref64):
            d = gm.diagnose(sb.BaroclinicQGState(q=sol.ys.q[i]))
            assert d.nonfinite == 0
            for key in ("kinetic_energy", "enstrophy"):
                got, want = np.asarray(getattr(d, key), np.float64), dref[key]
                floor = np.abs(ref32[i][key] - want) / np.abs(want)
                tol = np.maximum(1e-4, 3.0 * floor) if is32 else 1e-9
                assert np.all(np.abs(got - want) <= tol * np.abs(want)), (is32, i, key, got, want, floor)
    # shallow water
    osw, gsw = swm_pair(64, 64, np.float32, "periodic")
    h, u, v = swm_state(64, 64, np.float32, noise=False)
    dt, nstep, every = 20.0, 2000, 200
    ts = [every * dt * (i + 1) for i in range(nstep // every)]
    sol = gsw.integrate(sb.MultilayerSW2DState(h=h, u=u, v=v), 0.0, ts[-1], dt, saveat=sb.SaveAt(ts=ts), max_steps=None)
    series = []
    osw.integrate(*[a.astype(np.float64) for a in (h, u, v)], 0.0, ts[-1], dt,
                  on_step=lambda i, y: series.append(osw.diagnose(*y)) if i % every == 0 else None)
    for i, dref in enumerate(series):
        d = gsw.diagnose(sb.MultilayerSW2DState(h=sol.ys.h[i], u=sol.ys.u[i], v=sol.ys.v[i]))
        assert np.allclose(d.energy, dref["energy"], rtol=1e-4), i
        assert np.allclose(d.enstrophy, dref["enstrophy"], rtol=1e-4), i


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reparameterized_qg_parity(dtype):
    """ReparameterizedQG (qg/reparameterized.py:66-330): project, apply_boundary_conditions, the
    `_rhs` of build_terms, 1 and 20 Tsit5 steps and the diagnostics against the oracle; the reference's
    property tests (tests/models/test_qg_reparameterized.py) on the CUDA path."""
    import somax_b200 as sb
    from oracle.reparam import create_reparameterized
    kw = dict(nx=64, ny=48, lateral_viscosity=15.0, bottom_drag=3.6e-8, wind_amplitude=8e-5)
    om = create_reparameterized(**kw)
    gm = sb.ReparameterizedQG.create(dtype=np.dtype(dtype).name, **kw)
    assert isinstance(gm.swm, sb.MultilayerShallowWater2D)
    rng = np.random.default_rng(4)
    H = np.asarray(om.swm.H)[:, None, None]
    h = (H * (1.0 + 1e-3 * rng.standard_normal((3, 50, 66)))).astype(dtype)
    u = (0.05 * rng.standard_normal(h.shape)).astype(dtype)
    v = (0.05 * rng.standard_normal(h.shape)).astype(dtype)
    st = sb.MultilayerSW2DState(h=h, u=u, v=v)
    f64 = [a.astype(np.float64) for a in (h, u, v)]
    tol = 2e-5 if dtype == np.float32 else 1e-10
    for got, ref in ((gm.project(st), om.project(*f64)), (gm.apply_boundary_conditions(st), om.bc(*f64))):
        for name, r in zip("huv", ref):
            assert rel(getattr(got, name), r) <= tol, name
    # vector_field is the plain shallow-water one; _rhs evaluates it at the projected state
    t = gm.vector_field(0.0, st)
    for name, r, rs in zip("huv", om.rhs(*f64), om.rhs(h, u, v)):
        assert_swm_parity(getattr(t, name), r, rs, dtype, name)
    t = gm.build_terms().vf(0.0, st)
    for name, r in zip("huv", om.rhs(*om.bc(*f64))):
        assert rel(getattr(t, name), r) <= (1e-3 if dtype == np.float32 else 1e-9), name
    dt = 200.0
    for steps in (1, 20):
        sol = gm.integrate(st, 0.0, steps * dt, dt)
        ref = om.integrate(*f64, 0.0, steps * dt, dt)
        same = om.integrate(h, u, v, 0.0, steps * dt, dt) if dtype == np.float32 else ref
        for name, r, rs in zip("huv", ref, same):
            # fp32: the shallow-water momentum tendencies sit on a rounding floor (assert_swm_parity)
            lim = max(1e-5, 1.5 * rel(rs, r)) if dtype == np.float32 else 1e-10
            assert rel(getattr(sol.ys, name)[0], r) <= lim, (steps, name)
    last = sb.MultilayerSW2DState(h=sol.ys.h[0], u=sol.ys.u[0], v=sol.ys.v[0])
    d, dref = gm.diagnose(last), om.diagnose(*ref)
    assert isinstance(d, sb.ReparamQGDiagnostics)
    assert rel(d.psi, dref["psi"]) <= tol * 50
    assert rel(d.u_ageostrophic, dref["u_ageostrophic"]) <= tol * 500
    assert np.allclose(d.energy, dref["energy"], rtol=1e-4)
    # rest state: fixed point of the projection, zero tendency (test_qg_reparameterized.py:60-125)
    m2, st0 = sb.gfd_testcases.doublegyre_reparameterized_qg(nx=32, ny=32, wind_amplitude=0.0, dtype=np.dtype(dtype).name)
    p = m2.project(st0)
    assert np.allclose(p.h, st0.h, atol=1e-3 if dtype == np.float32 else 1e-6) and np.abs(p.u).max() < 1e-10
    tend = m2.build_terms().vf(0.0, st0)
    assert max(np.abs(tend.h).max(), np.abs(tend.u).max(), np.abs(tend.v).max()) < 1e-10
    with pytest.raises(ValueError, match="wall"):
        sb.ReparameterizedQG.create(nx=16, ny=16, bc="periodic")
