"""GPU: the somax-sim run layer (RunSpec -> chunked stepping on the device -> zarr artifacts).

Ports of the reference's tests/test_cli_run.py for the runner row of the hot-path table: artifact
set and attrs, save-time grid in the snapshot store, spinup -> restart continuity, the
`swm 64^2 dt=300 must diverge` behavioural pin (:187-212) and the finite 32^2 dt=10 run (:158-179).
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def qg_spec(t1=6 * 3600.0, si=2 * 3600.0, nx=64):
    from somax_b200.cli import RunSpec
    return RunSpec.from_dict({
        "testcase": {"name": "doublegyre_baroclinic_qg",
                     "grid": {"nx": nx, "ny": nx, "Lx": 4e6, "Ly": 4e6},
                     "consts": {"f0": 9.375e-5, "beta": 1.754e-11, "n_layers": 3},
                     "stratification": {"H": [400.0, 1100.0, 2600.0], "g_prime": [9.81, 0.025, 0.0125]},
                     "params": {"lateral_viscosity": 15.0, "bottom_drag": 1e-7, "wind_amplitude": 1.3e-10}},
        "timestepping": {"t0": 0.0, "t1": t1, "dt": 600.0, "save_interval": si}})


def swm_spec(nx, dt, t1, si):
    from somax_b200.cli import RunSpec
    return RunSpec.from_dict({
        "testcase": {"name": "baroclinic_instability_swm",
                     "grid": {"nx": nx, "ny": nx, "Lx": 1e6, "Ly": 1e6},
                     "consts": {"f0": 1e-4, "beta": 1.6e-11},
                     "stratification": {"H": [500.0, 4500.0], "g_prime": [9.81, 0.025]},
                     "params": {"lateral_viscosity": 100.0, "bottom_drag": 1e-7, "jet_speed": 0.5,
                                "jet_width": 5e4, "perturbation": 0.01}},
        "timestepping": {"t0": 0.0, "t1": t1, "dt": dt, "save_interval": si}})


def test_simulate_writes_artifacts_and_matches_direct_integration(tmp_path):
    import somax_b200 as sb
    from somax_b200 import io
    from somax_b200.cli import simulate
    spec = qg_spec()
    spec.validate()
    res = simulate(spec, tmp_path / "run", diagnostics_per_save=2)
    assert res.snapshots_path.name == "snapshots.zarr" and res.final_state_path.name == "final_state.zarr"
    assert res.n_steps == 36 and res.metrics_path.exists() and (tmp_path / "run" / "resolved.yaml").exists()
    snaps = io.load_dataset(res.snapshots_path)
    assert snaps["time"].values.tolist() == [0.0, 7200.0, 14400.0, 21600.0]
    assert snaps["q"].dims == ("time", "layer", "y", "x") and snaps["q"].shape == (4, 3, 66, 66)
    assert snaps.attrs["somax_sim_mode"] == "run" and snaps.attrs["testcase_name"] == "doublegyre_baroclinic_qg"
    assert snaps.attrs["state_class"] == "BaroclinicQGState" and snaps.attrs["dt"] == 600.0
    final = io.dataset_to_state(io.load_dataset(res.final_state_path))
    assert isinstance(final, sb.BaroclinicQGState)
    assert np.array_equal(final.q, snaps["q"].values[-1])
    # the chunked run = the same chunks integrated directly (BC of state0 re-applied per chunk, as
    # in the reference runner)
    model, st = sb.gfd_testcases.doublegyre_baroclinic_qg(nx=64, ny=64)
    q = st.q
    for _ in range(6):
        q = model.integrate(sb.BaroclinicQGState(q=q), 0.0, 3600.0, 600.0).ys.q[0]
    assert np.array_equal(final.q, q)
    assert np.isfinite(final.q).all() and np.abs(final.q).max() > 0
    metrics = json.loads(res.metrics_path.read_text())
    assert metrics["mode"] == "run" and metrics["n_steps"] == 36 and metrics["total_kinetic_energy"] > 0
    assert "kinetic_energy_layer_2" in metrics and "psi_max" in metrics
    log = (tmp_path / "run" / "run.log").read_text().splitlines()
    assert "| somax-sim/run    | started" in log[0]
    chunks = [ln for ln in log if "| chunk " in ln]
    assert len(chunks) == 7 and "chunk 0/6 sim_t=0 s | q[1/s]=[0,0,0]" in chunks[0] and "initial state" in chunks[0]
    assert "chunk 6/6 sim_t=2.16e+04 s (6.00 hr)" in chunks[-1] and "physics: total_kinetic_energy=" in chunks[-1]
    assert log[-1].endswith("finished cleanly")


def test_spinup_then_restart_continues(tmp_path):
    from somax_b200 import io
    from somax_b200.cli import restart, simulate, spinup
    sp = spinup(qg_spec(t1=4 * 3600.0, si=3600.0), tmp_path / "spin")
    assert sp.snapshots_path is None and sp.metrics_path is None and sp.final_state_path.exists()
    assert not (tmp_path / "spin" / "snapshots.zarr").exists()
    ds = io.load_dataset(sp.final_state_path)
    assert ds.attrs["somax_sim_mode"] == "spinup" and ds["time"].values.tolist() == [4 * 3600.0]
    rs = restart(qg_spec(t1=2 * 3600.0, si=3600.0), tmp_path / "prod", restart_from=sp.final_state_path)
    first = io.load_dataset(rs.snapshots_path)["q"].values[0]
    assert np.array_equal(first, ds["q"].values[0])            # the run starts from the stored state
    assert io.load_dataset(rs.final_state_path).attrs["somax_sim_mode"] == "restart"
    with pytest.raises(TypeError, match="expects MultilayerSW2DState"):
        restart(swm_spec(32, 10.0, 100.0, 50.0), tmp_path / "bad", restart_from=sp.final_state_path)


def test_swm_finite_run_and_divergence_guard(tmp_path):
    from somax_b200.cli import IntegrationDivergedError, simulate
    ok = simulate(swm_spec(32, 10.0, 3600.0, 1800.0), tmp_path / "ok")
    assert ok.n_steps == 360 and ok.snapshots_path.exists()
    # reference pin: the jet at 64^2 with dt = 300 s must blow up within 48 steps, and the runner
    # must refuse to write artifacts
    with pytest.raises(IntegrationDivergedError, match="non-finite values during chunk"):
        simulate(swm_spec(64, 300.0, 48 * 300.0, 12 * 300.0), tmp_path / "boom")
    out = tmp_path / "boom"
    assert not (out / "snapshots.zarr").exists() and not (out / "final_state.zarr").exists()
    assert not (out / "metrics.json").exists()
    log = (out / "run.log").read_text()
    assert "ABORT at chunk" in log and "FAILED during integrate: IntegrationDivergedError" in log


def test_cli_run_debug_config_and_assertions(tmp_path):
    """`somax-sim run --config configs/short/swm_jet.yaml --debug` end to end; a violated CFL assertion
    stops the run before any stepping."""
    import os
    from somax_b200 import io
    from somax_b200.cli import AssertionFailedError, app, load_yaml, simulate
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = os.path.join(root, "configs", "short", "swm_jet.yaml")
    assert app.main(["run", "--config", cfg, "--output-dir", str(tmp_path / "o"), "--debug",
                     "--diagnostics-per-save", "2"]) == 0
    snaps = io.load_dataset(tmp_path / "o" / "snapshots.zarr")
    assert snaps["h"].shape == (3, 2, 34, 34) and snaps["time"].values.tolist() == [0.0, 1800.0, 3600.0]
    assert np.isfinite(snaps["u"].values).all()
    spec = load_yaml(cfg).with_debug_applied()                    # (drops `assertions`, as the reference's does)
    spec.assertions = {"cfl": {"wave_speed_m_per_s": 221.5, "max_cfl": 0.5}}
    spec.timestepping.dt = 300.0                                  # gravity-wave CFL ~ 2
    with pytest.raises(AssertionFailedError, match="cfl check FAILED"):
        simulate(spec, tmp_path / "cfl")
    assert not (tmp_path / "cfl" / "run.log").exists()
    spec = load_yaml(cfg).with_debug_applied()
    spec.assertions = {"bounded_metric": {"name": "total_energy", "max": 1.0}}
    with pytest.raises(AssertionFailedError, match="above max"):
        simulate(spec, tmp_path / "post")
    assert not (tmp_path / "post" / "final_state.zarr").exists()
