"""CPU: RunSpec, the test-case registry, the save-time grid and the zarr v3 dataset IO.

Ports of the reference's host-side tests for the "next" rows of the hot-path table
(tests/test_io_xarray.py, tests/test_cli_spec.py, tests/test_cli_run.py:341-381); nothing here
computes model physics.
"""
import json
from dataclasses import dataclass

import numpy as np
import pytest

from somax_b200 import io
from somax_b200.cli import spec as S
from somax_b200.cli._run import (_build_diagnostic_grid, _build_save_times, _flatten_diagnostics,
                                 format_field_stats, format_time_seconds, format_wallclock)
from somax_b200.core import State
from somax_b200.models.qg import BaroclinicQGState, BarotropicQGState
from somax_b200.models.swm import MultilayerSW2DState


def make_spec(**ts):
    t = dict(t0=0.0, t1=86400.0, dt=600.0, save_interval=21600.0)
    t.update(ts)
    return S.RunSpec(testcase=S.TestCaseSpec("doublegyre_qg", grid={"nx": 64, "ny": 64, "Lx": 1e6, "Ly": 1e6},
                                             consts={"f0": 1e-4, "beta": 1.6e-11},
                                             params={"lateral_viscosity": 500.0, "bottom_drag": 1e-7,
                                                     "wind_amplitude": 1e-12}),
                     timestepping=S.TimesteppingSpec(**t))


# ------------------------------------------------------------------ spec
def test_spec_validate_and_errors():
    make_spec().validate()
    for bad, msg in ((dict(t1=0.0), "must be >"), (dict(dt=0.0), "dt"), (dict(save_interval=0.0), "save_interval"),
                     (dict(save_interval=1e9), "cannot exceed")):
        with pytest.raises(ValueError, match=msg):
            make_spec(**bad).validate()
    sp = make_spec()
    sp.testcase.name = ""
    with pytest.raises(ValueError, match="non-empty"):
        sp.validate()


def test_spec_debug_merge_is_deep_and_consumed():
    sp = make_spec()
    assert sp.with_debug_applied() is sp
    sp.debug = S.DebugSpec(testcase={"grid": {"nx": 32}}, timestepping={"t1": 3600.0})
    out = sp.with_debug_applied()
    assert out.testcase.grid == {"nx": 32, "ny": 64, "Lx": 1e6, "Ly": 1e6}
    assert sp.testcase.grid["nx"] == 64                       # original untouched
    assert out.timestepping.t1 == 3600.0 and out.timestepping.dt == 600.0
    assert out.debug == S.DebugSpec()
    sp.assertions = {"cfl": {"wave_speed_m_per_s": 1.0}}
    assert sp.with_debug_applied().assertions == {}           # reference behaviour (spec.py:225-230): not carried over
    sp.debug = S.DebugSpec(testcase={"nope": {}})
    with pytest.raises(ValueError, match="does not match"):
        sp.with_debug_applied()
    sp.debug = S.DebugSpec(testcase={"grid": 3})
    with pytest.raises(ValueError, match="both sides"):
        sp.with_debug_applied()


def test_spec_yaml_round_trip(tmp_path):
    sp = make_spec()
    sp.assertions = {"cfl": {"max": 0.5}}
    p = tmp_path / "c.yaml"
    S.dump_yaml(sp, str(p))
    back = S.load_yaml(str(p))
    assert back.to_dict() == sp.to_dict()
    with pytest.raises(ValueError, match="missing required block"):
        S.RunSpec.from_dict({"testcase": {"name": "x"}})
    (tmp_path / "l.yaml").write_text("- 1\n- 2\n")
    with pytest.raises(ValueError, match="top-level mapping"):
        S.load_yaml(str(tmp_path / "l.yaml"))


def test_registry_names():
    from somax_b200.cli import get_adapter, list_test_cases
    assert list_test_cases() == ["baroclinic_instability_swm", "barotropic_jet_instability",
                                 "doublegyre_baroclinic_qg", "doublegyre_qg"]
    with pytest.raises(KeyError, match="unknown test case"):
        get_adapter("nope")


# ------------------------------------------------------------------ save grid (tests/test_cli_run.py:341-381)
def test_save_times_exact_multiple_and_remainder():
    ts = _build_save_times(make_spec(t1=86400.0, save_interval=21600.0))
    assert ts.dtype == np.float64 and np.allclose(ts, [0, 21600, 43200, 64800, 86400])
    ts = _build_save_times(make_spec(t1=100.0, dt=1.0, save_interval=30.0))
    assert np.allclose(ts, [0, 30, 60, 90, 100])               # last interval shorter, t1 appended
    assert np.allclose(_build_save_times(make_spec(), only_final=True), [0, 86400.0])
    ts = _build_save_times(make_spec(t0=10.0, t1=20.0, dt=1.0, save_interval=10.0))
    assert np.allclose(ts, [10.0, 20.0])


def test_diagnostic_grid_subdivides():
    save = np.asarray([0.0, 10.0, 20.0])
    g, idx = _build_diagnostic_grid(save, 1)
    assert g is save and idx == {0, 1, 2}
    g, idx = _build_diagnostic_grid(save, 4)
    assert np.allclose(g, np.linspace(0, 20, 9)) and idx == {0, 4, 8}
    assert np.allclose(g[sorted(idx)], save)


def test_formatters():
    assert format_time_seconds(42) == "42 s"
    assert format_time_seconds(600) == "600 s (10.00 min)"
    assert format_time_seconds(86400) == "8.64e+04 s (1.00 day)"
    assert format_wallclock(0.5) == "500 ms" and format_wallclock(2.0) == "2.00 s"
    assert format_wallclock(125.0) == "2m5s" and format_wallclock(3725.0) == "1h2m5s"
    assert format_field_stats("h", min_val=1, mean_val=2, max_val=3, nan_count=0) == "h[m]=[1,2,3]"
    assert format_field_stats("zz", min_val=1, mean_val=2, max_val=3, nan_count=2) == "zz=[1,2,3] NaN=2"


def test_flatten_diagnostics():
    @dataclass
    class D:
        a: object
        b: object
        c: object
        d: object = None
    flat = _flatten_diagnostics(D(a=np.float32(2.0), b=np.array([1.0, 3.0]), c=np.arange(40.0).reshape(5, 8)))
    assert flat == {"a": 2.0, "b_layer_0": 1.0, "b_layer_1": 3.0, "c_mean": 19.5, "c_max": 39.0, "c_min": 0.0}
    assert _flatten_diagnostics(None) == {} and _flatten_diagnostics(3) == {}


# ------------------------------------------------------------------ datasets (tests/test_io_xarray.py)
def test_dims_by_rank_and_time_axis():
    q = np.arange(3 * 6 * 5, dtype=np.float32).reshape(3, 6, 5)
    ds = io.state_to_dataset(BaroclinicQGState(q=q))
    assert ds["q"].dims == ("layer", "y", "x") and "time" not in ds.coords
    ds = io.state_to_dataset(BarotropicQGState(q=q[0]), time=12.5)
    assert ds["q"].dims == ("time", "y", "x") and ds["q"].shape == (1, 6, 5)
    assert ds["time"].values.dtype == np.float64 and ds["time"].values.tolist() == [12.5]
    assert ds.attrs["state_class"] == "BarotropicQGState" and ds.attrs["state_module"] == "somax_b200.models.qg"

    @dataclass
    class Odd(State):
        x: object
        w: object
    ds = io.state_to_dataset(Odd(x=np.zeros(4), w=np.zeros((2, 2, 2, 2))))
    assert ds["x"].dims == ("x",) and ds["w"].dims == ("dim0", "dim1", "dim2", "dim3")


def test_snapshots_errors_and_round_trip():
    h = np.random.default_rng(0).standard_normal((5, 2, 4, 3)).astype(np.float32)
    snaps = MultilayerSW2DState(h=h, u=h + 1, v=h + 2)
    with pytest.raises(ValueError, match="leading dim 5 but ts has length 3"):
        io.snapshots_to_dataset(snaps, np.arange(3.0))
    with pytest.raises(ValueError, match="ts must be 1-D"):
        io.snapshots_to_dataset(snaps, np.zeros((5, 1)))
    ds = io.snapshots_to_dataset(snaps, np.arange(5.0), attrs={"dt": 2.0})
    assert ds["u"].dims == ("time", "layer", "y", "x") and ds.attrs["dt"] == 2.0
    last = io.dataset_to_state(ds)                               # default: time_index = -1
    assert isinstance(last, MultilayerSW2DState) and np.array_equal(last.v, h[-1] + 2)
    assert np.array_equal(io.dataset_to_state(ds, time_index=1).h, h[1])


def test_dataset_to_state_errors():
    ds = io.state_to_dataset(BarotropicQGState(q=np.zeros((4, 4), np.float32)))
    only_h = io.Dataset({"h": (("layer", "y", "x"), np.zeros((2, 4, 4), np.float32))})
    with pytest.raises(ValueError, match="missing variable 'u'"):
        io.dataset_to_state(only_h, MultilayerSW2DState)
    ds.attrs.pop("state_class")
    with pytest.raises(ValueError, match="state_class"):
        io.dataset_to_state(ds)
    ds.attrs.update(state_class="Popen", state_module="subprocess")
    with pytest.raises(ValueError, match="refusing to auto-import"):
        io.dataset_to_state(ds)
    # a store written by the reference itself names its own module tree: mapped to this package
    ds.attrs.update(state_class="BarotropicQGState", state_module="somax._src.models.qg.barotropic")
    assert isinstance(io.dataset_to_state(ds), BarotropicQGState)


def test_zarr_v3_store_layout_round_trip_and_append(tmp_path):
    rng = np.random.default_rng(1)
    q = rng.standard_normal((3, 3, 6, 5)).astype(np.float32)
    ds = io.snapshots_to_dataset(BaroclinicQGState(q=q), np.asarray([0.0, 10.0, 20.0]),
                                 attrs={"testcase_name": "t", "dt": np.float64(0.5)})
    store = tmp_path / "snapshots.zarr"
    io.save_dataset(ds, store, mode="w")
    root = json.loads((store / "zarr.json").read_text())
    assert root["zarr_format"] == 3 and root["node_type"] == "group"
    assert root["attributes"]["state_class"] == "BaroclinicQGState" and root["attributes"]["dt"] == 0.5
    meta = json.loads((store / "q" / "zarr.json").read_text())
    assert meta["shape"] == [3, 3, 6, 5] and meta["data_type"] == "float32"
    assert meta["dimension_names"] == ["time", "layer", "y", "x"]
    assert meta["chunk_grid"]["configuration"]["chunk_shape"] == [1, 3, 6, 5]
    assert [c["name"] for c in meta["codecs"]] == ["bytes"]
    assert (store / "q" / "c" / "2" / "0" / "0" / "0").stat().st_size == 3 * 6 * 5 * 4
    back = io.load_dataset(store)
    assert np.array_equal(back["q"].values, q) and back["time"].values.tolist() == [0.0, 10.0, 20.0]
    assert back.attrs["state_module"] == "somax_b200.models.qg"       # attrs survive the round trip
    assert np.array_equal(io.dataset_to_state(back).q, q[-1])
    with pytest.raises(FileExistsError):
        io.save_dataset(ds, store, mode="w-")
    more = io.snapshots_to_dataset(BaroclinicQGState(q=q[:2] + 7), np.asarray([30.0, 40.0]))
    io.append_to_dataset(more, store)
    back = io.load_dataset(store)
    assert back["q"].shape == (5, 3, 6, 5) and np.array_equal(back["q"].values[3:], q[:2] + 7)
    assert back["time"].values.tolist() == [0.0, 10.0, 20.0, 30.0, 40.0]
    with pytest.raises(ValueError, match="differ"):
        io.append_to_dataset(io.snapshots_to_dataset(BaroclinicQGState(q=q[:1, :, :3]), np.asarray([50.0])), store)
    io.save_dataset(io.state_to_dataset(BaroclinicQGState(q=q[0]), time=1.0), store, mode="w")   # overwrite
    assert io.load_dataset(store)["q"].shape == (1, 3, 6, 5)


def test_async_snapshot_writer_host_states(tmp_path):
    w = io.AsyncSnapshotWriter(tmp_path / "s.zarr", BarotropicQGState, attrs={"k": 1})
    for t in range(4):
        w.put(float(t), BarotropicQGState(q=np.full((4, 3), t, np.float64)))
    w.close()
    ds = io.load_dataset(tmp_path / "s.zarr")
    assert ds["q"].shape == (4, 4, 3) and ds["q"].values[3].max() == 3.0 and ds.attrs["k"] == 1
    assert ds["time"].values.tolist() == [0.0, 1.0, 2.0, 3.0]


# ------------------------------------------------------------------ assertions + CLI (host only)
def test_assertions_cfl_and_bounded_metric():
    from types import SimpleNamespace
    from somax_b200.cli._assertions import (AssertionFailedError, check_bounded_metric, check_cfl, run_postflight,
                                            run_preflight)
    sp = make_spec()                                              # dt = 600
    model = SimpleNamespace(grid=SimpleNamespace(dx=15625.0, dy=15625.0))
    check_cfl(sp, model, wave_speed_m_per_s=1.0)
    with pytest.raises(AssertionFailedError, match="CFL = 8.506 > max_cfl = 0.5"):
        check_cfl(sp, model, wave_speed_m_per_s=221.5)
    with pytest.raises(AssertionFailedError, match="no .grid"):
        check_cfl(sp, SimpleNamespace(), wave_speed_m_per_s=1.0)
    check_bounded_metric(sp, {"e": 2.0}, name="e", min=1.0, max=3.0)
    for metrics, kw, msg in (({}, {}, "not present"), ({"e": float("nan")}, {}, "non-finite"),
                             ({"e": 0.5}, {"min": 1.0}, "below min"), ({"e": 5.0}, {"max": 3.0}, "above max"),
                             ({"e": "x"}, {}, "not a numeric")):
        with pytest.raises(AssertionFailedError, match=msg):
            check_bounded_metric(sp, metrics, name="e", **kw)
    sp.assertions = {"cfl": {"wave_speed_m_per_s": 1.0}, "bounded_metric": {"name": "e", "max": 1.0}}
    run_preflight(sp, model)                                      # postflight names are skipped
    with pytest.raises(AssertionFailedError, match="above max"):
        run_postflight(sp, {"e": 2.0})
    sp.assertions = {"typo": {}}
    with pytest.raises(AssertionFailedError, match="unknown assertion 'typo'"):
        run_preflight(sp, model)


def test_cli_discovery_and_shipped_configs(capsys):
    import glob
    import os
    from somax_b200.cli import app
    assert app.main(["list-testcases"]) == 0
    assert "  - doublegyre_baroclinic_qg" in capsys.readouterr().out
    assert app.main(["list-models"]) == 0
    out = capsys.readouterr().out
    assert "  - BaroclinicQG" in out and "  - MultilayerShallowWater2D" in out
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfgs = sorted(glob.glob(os.path.join(root, "configs", "simulation", "*.yaml")) +
                  glob.glob(os.path.join(root, "configs", "short", "*.yaml")))
    assert len(cfgs) >= 8
    for c in cfgs:
        sp = S.load_yaml(c)
        dbg = sp.with_debug_applied()
        dbg.validate()
        assert dbg.testcase.grid["nx"] < sp.testcase.grid["nx"]
        assert app.main(["show-config", c]) == 0
    assert "# Resolved RunSpec from" in capsys.readouterr().out


# ------------------------------------------------------------------ chunk loop with a stand-in model (host logic only)
class _DecayModel:
    """dq/dt = -k q integrated exactly: a stand-in with the three methods the chunk loop calls."""

    def __init__(self, blow_up_after=None):
        self.k, self.calls, self.blow_up_after = 1e-3, [], blow_up_after

    def integrate(self, state, t0, t1, dt, max_steps=None):
        from types import SimpleNamespace
        self.calls.append((t0, t1, dt, max_steps))
        q = np.asarray(state.q) * np.exp(-self.k * (t1 - t0))
        if self.blow_up_after is not None and len(self.calls) > self.blow_up_after:
            q = q.copy()
            q[0, 0] = np.nan
        return SimpleNamespace(ys=BarotropicQGState(q=q[None]), stats={"num_steps": int(np.ceil((t1 - t0) / dt))})

    def diag_scalars(self, state):
        q = np.nan_to_num(np.asarray(state.q))
        return np.array([float((q ** 2).sum())]), np.array([float(np.abs(q).sum())]), 0.0


def test_chunk_loop_logs_snapshots_and_aborts(tmp_path):
    from somax_b200.cli._run import IntegrationDivergedError, RunLog, _chunked_integrate_with_diagnostics
    save_ts = np.asarray([0.0, 100.0, 200.0, 250.0])
    state0 = BarotropicQGState(q=np.ones((4, 5)))
    model, snaps = _DecayModel(), []
    log = RunLog(tmp_path, label="somax-sim/run")
    final, n_steps = _chunked_integrate_with_diagnostics(
        model, state0, save_ts, 10.0, diagnostics_per_save=2, max_steps_per_chunk=77, run_log=log, mode="run",
        on_snapshot=lambda i, t, st: snaps.append((i, t, float(np.asarray(st.q)[0, 0]))))
    log.stop("finished cleanly")
    # 3 save intervals x 2 sub-chunks, each integrated with the per-chunk step cap
    assert [(a, b) for a, b, _, _ in model.calls] == [(0, 50), (50, 100), (100, 150), (150, 200), (200, 225), (225, 250)]
    assert all(c[3] == 77 for c in model.calls) and n_steps == 26
    assert [s[:2] for s in snaps] == [(0, 0.0), (1, 100.0), (2, 200.0), (3, 250.0)]     # snapshots only at save times
    assert np.isclose(snaps[-1][2], np.exp(-0.25)) and np.isclose(np.asarray(final.q)[0, 0], np.exp(-0.25))
    lines = (tmp_path / "run.log").read_text().splitlines()
    chunk = [ln for ln in lines if "| chunk " in ln]
    assert len(chunk) == 7 and "chunk 0/6 sim_t=0 s | q[1/s]=[1,1,1] | physics: kinetic_energy=20 enstrophy=20 | initial state" in chunk[0]
    assert "chunk 6/6 sim_t=250 s (4.17 min) | q[1/s]=[0.779,0.779,0.779]" in chunk[-1] and "| wall=" in chunk[-1]
    assert any("all 6 chunks completed in" in ln for ln in lines) and lines[-1].endswith("finished cleanly")

    model, snaps = _DecayModel(blow_up_after=2), []
    (tmp_path / "b").mkdir()
    log = RunLog(tmp_path / "b", label="somax-sim/run")
    with pytest.raises(IntegrationDivergedError, match=r"during chunk 3/6 \(sim_t=150 s \(2.50 min\)\)"):
        _chunked_integrate_with_diagnostics(model, state0, save_ts, 10.0, diagnostics_per_save=2,
                                            max_steps_per_chunk=77, run_log=log, mode="run",
                                            on_snapshot=lambda i, t, st: snaps.append(i))
    log.stop("FAILED")
    assert len(model.calls) == 3 and snaps == [0, 1]            # stops at the first non-finite chunk
    text = (tmp_path / "b" / "run.log").read_text()
    assert "NaN=1" in text and "ABORT at chunk 3/6: non-finite state" in text


def test_energy_growth_warning(tmp_path):
    from types import SimpleNamespace
    from somax_b200.cli._run import RunLog, _chunked_integrate_with_diagnostics

    class Growing(_DecayModel):
        def integrate(self, state, t0, t1, dt, max_steps=None):
            return SimpleNamespace(ys=BarotropicQGState(q=(np.asarray(state.q) * 5.0)[None]), stats={"num_steps": 1})

    log = RunLog(tmp_path, label="somax-sim/run")
    _chunked_integrate_with_diagnostics(Growing(), BarotropicQGState(q=np.ones((3, 3))), np.asarray([0.0, 1.0, 2.0]), 1.0,
                                        diagnostics_per_save=1, max_steps_per_chunk=10, run_log=log, mode="run")
    log.stop()
    assert "!! energy grew 25.0x from previous chunk" in (tmp_path / "run.log").read_text()


# ------------------------------------------------------------------ more ports of the reference's host-side tests
def test_allowlist_prefix_and_explicit_class(tmp_path):
    """tests/test_io_xarray.py:286-306: 'somaxxx' is not under 'somax.'; an explicit state_class
    skips the allowlist and fails later on the missing variables."""
    from somax_b200.models.swm import NonlinearSW2DState
    for mod in ("os", "somaxxx.evil", "somax_b200x.evil"):
        ds = io.Dataset({"h": (("y", "x"), np.ones((4, 4), np.float32))}, attrs={"state_class": "Whatever", "state_module": mod})
        io.save_dataset(ds, tmp_path / "t.zarr")
        loaded = io.load_dataset(tmp_path / "t.zarr")
        with pytest.raises(ValueError, match="refusing to auto-import"):
            io.dataset_to_state(loaded)
        with pytest.raises(ValueError, match="missing variable"):
            io.dataset_to_state(loaded, state_class=NonlinearSW2DState)


def test_assert_finite_state_messages():
    """tests/test_cli_run.py:51-111."""
    from somax_b200.cli._run import IntegrationDivergedError, _assert_finite_state
    from somax_b200.models.swm import NonlinearSW2DState
    z = np.zeros((3, 4, 4))
    _assert_finite_state(NonlinearSW2DState(h=np.ones((3, 4, 4)), u=z, v=z), mode="run")
    h = np.ones((3, 4, 4))
    h[1, 0, 0] = np.nan
    with pytest.raises(IntegrationDivergedError) as ei:
        _assert_finite_state(NonlinearSW2DState(h=h, u=z, v=z), mode="run")
    assert "h (1/48 non-finite)" in str(ei.value) and "Mid-integration blow-up" in str(ei.value)
    with pytest.raises(IntegrationDivergedError, match="non-finite"):
        _assert_finite_state(NonlinearSW2DState(h=np.full((2, 3, 3), np.inf), u=np.zeros((2, 3, 3)), v=np.zeros((2, 3, 3))), mode="run")
    h = np.ones((3, 4, 4))
    h[0] = np.nan
    with pytest.raises(IntegrationDivergedError) as ei:
        _assert_finite_state(NonlinearSW2DState(h=h, u=z, v=z), mode="run")
    assert "already non-finite at the first saved step" in str(ei.value) and "CFL" in str(ei.value)
    with pytest.raises(IntegrationDivergedError, match="spinup integration"):
        _assert_finite_state(NonlinearSW2DState(h=np.full((2, 3, 3), np.nan), u=np.zeros((2, 3, 3)), v=np.zeros((2, 3, 3))), mode="spinup")


def test_run_log_lifecycle(tmp_path):
    """tests/test_cli_progress.py:32-165."""
    import time as _time
    from somax_b200.cli._run import RUN_LOG_FORMAT, start_run_log, stop_run_log
    ctx = start_run_log(tmp_path / "deep" / "nested", label="test/run", alive_interval=10.0, enable_alive_thread=False)
    log = tmp_path / "deep" / "nested" / "run.log"
    assert log.exists() and "started" in log.read_text() and "test/run" in log.read_text()
    ctx.debug("chunk 1/3 | physics: x")
    stop_run_log(ctx, final_message="finished cleanly")
    stop_run_log(ctx)                                    # idempotent
    lines = log.read_text().splitlines()
    assert lines[-1].endswith("finished cleanly") and all("| test/run" in ln and "| DEBUG   |" in ln for ln in lines)
    assert len(lines) == 3 and "{message}" in RUN_LOG_FORMAT
    for bad in (0.0, -1.0):
        with pytest.raises(ValueError, match="alive_interval must be > 0"):
            start_run_log(tmp_path, label="x", alive_interval=bad)
    ctx = start_run_log(tmp_path / "alive", label="x", alive_interval=0.05)
    _time.sleep(0.3)
    stop_run_log(ctx)
    assert (tmp_path / "alive" / "run.log").read_text().count("alive (") >= 2
    ctx = start_run_log(tmp_path / "quiet", label="x", alive_interval=0.05, enable_alive_thread=False)
    _time.sleep(0.2)
    stop_run_log(ctx)
    assert "alive (" not in (tmp_path / "quiet" / "run.log").read_text()


def test_save_times_monthly_over_a_year_and_attrs():
    """tests/test_cli_run.py:362-412."""
    from somax_b200.cli._run import _attrs_for
    month, year = 2_592_000.0, 31_557_600.0
    sp = make_spec(t1=year, dt=600.0, save_interval=month)
    ts = _build_save_times(sp)
    assert len(ts) == 14 and np.allclose(np.diff(ts)[:12], month) and ts[-1] == year and ts[-1] - ts[-2] < month
    attrs = _attrs_for(sp, mode="run")
    assert attrs["somax_sim_mode"] == "run" and attrs["testcase_name"] == "doublegyre_qg"
    assert all(type(attrs[k]) is float for k in ("t0", "t1", "dt", "save_interval"))


def test_spinup_config_disables_outputs_and_production_enables():
    """tests/test_configs.py:74-90."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spin = S.load_yaml(os.path.join(root, "configs", "simulation", "spinup_bc_qg.yaml"))
    assert spin.output.write_snapshots is False and spin.output.write_metrics is False
    for name in ("doublegyre_bt_qg", "doublegyre_bc_qg", "swm_jet"):
        sp = S.load_yaml(os.path.join(root, "configs", "simulation", name + ".yaml"))
        assert sp.output.write_snapshots and sp.output.write_metrics
        assert S.RunSpec.from_dict(sp.to_dict()).to_dict() == sp.to_dict()


def test_shipped_configs_carry_the_reference_run_parameters():
    """configs/simulation/*.yaml are the reference's runs (configs/_authoring/*.py of the reference):
    the values that set the physics and the run length are pinned here."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    load = lambda n: S.load_yaml(os.path.join(root, "configs", "simulation", n + ".yaml"))  # noqa: E731
    jet = load("swm_jet")
    assert jet.testcase.params["lateral_viscosity"] == 1000.0 and jet.timestepping.t1 == 2592000.0
    assert jet.assertions["cfl"]["wave_speed_m_per_s"] == 221.0
    bt = load("doublegyre_bt_qg")
    assert bt.timestepping.t1 == 31557600.0 and bt.assertions["bounded_metric"]["name"] == "kinetic_energy"
    bc = load("doublegyre_bc_qg")
    assert bc.timestepping.t1 == 31557600.0 and bc.testcase.grid["nx"] == 128
    assert load("spinup_bc_qg").timestepping.t1 == 94672800.0


def test_reader_decodes_zstd_and_gzip_chunks(tmp_path):
    """A store as zarr-python writes it by default (zstd after the bytes codec) - what the reference's
    `ds.to_zarr(path, zarr_format=3)` produces (io/xarray.py:235-249) - is readable for restart."""
    import gzip
    import json
    import pyarrow as pa
    from somax_b200 import io
    q = np.arange(2 * 3 * 6 * 6, dtype=np.float32).reshape(2, 3, 6, 6)
    ds = io.Dataset({"q": (("time", "layer", "y", "x"), q)}, {"time": (("time",), np.array([0.0, 600.0]))},
                    {"state_class": "BaroclinicQGState", "state_module": "somax_b200.models.qg"})
    for codec, enc in (("zstd", lambda b: pa.compress(b, codec="zstd", asbytes=True)), ("gzip", gzip.compress)):
        store = tmp_path / f"s_{codec}.zarr"
        io.save_dataset(ds, store)
        for adir in (store / "q", store / "time"):
            m = json.loads((adir / "zarr.json").read_text())
            m["codecs"].append({"name": codec, "configuration": {"level": 0}})
            (adir / "zarr.json").write_text(json.dumps(m))
            for f in adir.rglob("*"):
                if f.is_file() and f.name != "zarr.json":
                    f.write_bytes(enc(f.read_bytes()))
        back = io.load_dataset(store)
        assert np.array_equal(back["q"].values, q) and back["time"].values.tolist() == [0.0, 600.0]
    m = json.loads((store / "q" / "zarr.json").read_text())
    m["codecs"][-1]["name"] = "blosc"
    (store / "q" / "zarr.json").write_text(json.dumps(m))
    with pytest.raises(ValueError, match="compressors"):
        io.load_dataset(store)
