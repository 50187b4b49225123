"""Chunked runner: integrate a RunSpec, log physical diagnostics between chunks, write artifacts.

Mirror of somax/_src/cli/_run.py (`simulate` :783-807, `spinup` :810-832, `restart` :835-878,
`_integrate_and_write` :202-396, `_chunked_integrate_with_diagnostics` :540-700): same save-time
grid, diagnostic sub-chunks, `run.log` line layout, artifact set and failure semantics.

B200-first differences in HOW: the state stays resident on the device between chunks (CUDA
tensors in, CUDA tensors out of `model.integrate`); the per-chunk report needs only scalars,
which come from the library's fused diagnostics reduction (`diag_scalars`: energies, enstrophies
and the non-finite count in one pass) plus device min/mean/max reductions - a few hundred bytes
cross PCIe per chunk.  Full states leave the device only at snapshot times, through
`io.AsyncSnapshotWriter` (side-stream D2H into pinned buffers, zarr chunks written by a worker
thread) so stepping continues while a snapshot drains.
"""
from __future__ import annotations

import dataclasses
import json
import logging
import threading
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from .. import io
from . import _assertions
from ._factories import get_adapter
from .spec import RunSpec, dump_yaml

logger = logging.getLogger("somax_b200.sim")


class IntegrationDivergedError(RuntimeError):
    """The integration produced non-finite values; no artifacts are written (cli/_run.py:40-52)."""


@dataclass
class SimulationResult:
    output_dir: Path
    snapshots_path: Path | None
    final_state_path: Path
    metrics_path: Path | None
    wallclock_seconds: float
    n_steps: int | None


# ----------------------------------------------------------------------------------------
# formatting (cli/_units.py)
# ----------------------------------------------------------------------------------------
FIELD_UNITS = {"h": "m", "u": "m/s", "v": "m/s", "q": "1/s", "psi": "m²/s"}
_TIME_UNITS = [(3600.0, "min", 60.0), (86400.0, "hr", 3600.0), (2_592_000.0, "day", 86400.0),
               (31_557_600.0, "month", 2_592_000.0), (float("inf"), "yr", 31_557_600.0)]


def format_time_seconds(seconds: float) -> str:
    if seconds < 60.0:
        return f"{int(seconds)} s" if seconds == int(seconds) else f"{seconds:.3g} s"
    for cutoff, label, div in _TIME_UNITS:
        if seconds < cutoff:
            return f"{seconds:.3g} s ({seconds / div:.2f} {label})"
    return f"{seconds:.3g} s"   # pragma: no cover


def format_wallclock(seconds: float) -> str:
    if seconds < 1.0:
        return f"{seconds * 1000:.0f} ms"
    if seconds < 60.0:
        return f"{seconds:.2f} s"
    minutes, secs = divmod(seconds, 60.0)
    if minutes < 60:
        return f"{int(minutes)}m{secs:.0f}s"
    hours, mins = divmod(minutes, 60)
    return f"{int(hours)}h{int(mins)}m{secs:.0f}s"


def format_field_stats(name, *, min_val, mean_val, max_val, nan_count) -> str:
    unit = FIELD_UNITS.get(name, "")
    body = f"[{min_val:.3g},{mean_val:.3g},{max_val:.3g}]" + (f" NaN={nan_count}" if nan_count else "")
    return f"{name}{'[' + unit + ']' if unit else ''}={body}"


# ----------------------------------------------------------------------------------------
# run.log (cli/_progress.py): "<time> | <level> | <label> | <message>", truncated per run, plus a
# daemon thread that writes an "alive" line every 10 s
# ----------------------------------------------------------------------------------------
RUN_LOG_FORMAT = "{time:YYYY-MM-DD HH:mm:ss} | {level: <7} | {label: <16} | {message}"


class RunLog:
    """`<output_dir>/run.log` (truncated per run) + the periodic "alive" tick (cli/_progress.py:78-201)."""

    def __init__(self, output_dir, label: str, alive_interval: float = 10.0, enable_alive_thread: bool = True):
        if alive_interval <= 0:
            raise ValueError(f"alive_interval must be > 0 (got {alive_interval})")
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        self.path = out / "run.log"
        self.label = label
        self._f = open(self.path, "w")
        self._lock = threading.Lock()
        self._t0 = time.perf_counter()
        self._halt = threading.Event()
        self.debug("started")
        self._th = None
        if enable_alive_thread:
            self._th = threading.Thread(target=self._alive, args=(alive_interval,), daemon=True)
            self._th.start()

    def debug(self, msg: str):
        line = f"{time.strftime('%Y-%m-%d %H:%M:%S')} | {'DEBUG': <7} | {self.label: <16} | {msg}"
        with self._lock:
            if not self._f.closed:
                self._f.write(line + "\n")
                self._f.flush()
        logger.debug(msg)

    def _alive(self, interval):
        while not self._halt.wait(interval):
            self.debug(f"alive ({format_wallclock(time.perf_counter() - self._t0)} elapsed)")

    def stop(self, final_message: str | None = None):
        """Idempotent: a second call neither raises nor writes."""
        self._halt.set()
        if final_message is not None:
            self.debug(final_message)
        with self._lock:
            self._f.close()


def start_run_log(output_dir, *, label: str, alive_interval: float = 10.0, enable_alive_thread: bool = True) -> RunLog:
    return RunLog(output_dir, label, alive_interval, enable_alive_thread)


def stop_run_log(ctx: RunLog, *, final_message: str | None = None) -> None:
    ctx.stop(final_message)


# ----------------------------------------------------------------------------------------
# save grid and diagnostics grid (cli/_run.py:86-130, 487-537)
# ----------------------------------------------------------------------------------------
def _build_save_times(spec: RunSpec, *, only_final: bool = False) -> np.ndarray:
    """[t0, t0+SI, t0+2SI, ... (< t1), t1]; [t0, t1] for spinup.  float64."""
    ts = spec.timestepping
    t0, t1, si = float(ts.t0), float(ts.t1), float(ts.save_interval)
    if only_final:
        return np.asarray([t0, t1], dtype=np.float64)
    n_full = int((t1 - t0) // si)
    pts = [t0 + i * si for i in range(n_full + 1)]
    if abs(pts[-1] - t1) > 1e-9 * max(abs(t1), 1.0):
        pts.append(t1)
    return np.asarray(pts, dtype=np.float64)


def _build_diagnostic_grid(save_ts: np.ndarray, diagnostics_per_save: int):
    n = max(int(diagnostics_per_save), 1)
    if n == 1:
        return save_ts, set(range(save_ts.shape[0]))
    pts = [float(save_ts[0])]
    for i in range(save_ts.shape[0] - 1):
        pts.extend(float(x) for x in np.linspace(save_ts[i], save_ts[i + 1], n + 1)[1:])
    return np.asarray(pts), {i * n for i in range(save_ts.shape[0])}


# ----------------------------------------------------------------------------------------
# diagnostics
# ----------------------------------------------------------------------------------------
def _flatten_diagnostics(diag: Any) -> dict[str, Any]:
    """Scalars of a Diagnostics dataclass as a flat JSON-friendly dict (cli/_run.py:138-168)."""
    out: dict[str, Any] = {}
    if diag is None or not dataclasses.is_dataclass(diag):
        return out
    for f in dataclasses.fields(diag):
        v = getattr(diag, f.name)
        if v is None:
            continue
        try:
            a = io._host(v)
        except Exception:
            continue
        if a.ndim == 0:
            out[f.name] = float(a)
        elif a.ndim == 1 and a.size <= 16:
            for i, s in enumerate(a.tolist()):
                out[f"{f.name}_layer_{i}"] = float(s)
        else:
            out[f"{f.name}_mean"] = float(a.mean())
            out[f"{f.name}_max"] = float(a.max())
            out[f"{f.name}_min"] = float(a.min())
    return out


def _state_diagnostics(state) -> dict[str, dict[str, float]]:
    """Per-field min / mean / max over finite entries and the non-finite count (cli/_run.py:409-441).
    CUDA leaves are reduced on the device; four scalars per field reach the host."""
    out = {}
    for f in dataclasses.fields(state):
        leaf = getattr(state, f.name)
        if hasattr(leaf, "is_cuda"):
            import torch
            fin = torch.isfinite(leaf)
            n_bad = int(leaf.numel() - int(fin.sum()))
            if n_bad == leaf.numel():
                stats = (float("nan"),) * 3
            elif n_bad == 0:
                stats = (float(leaf.min()), float(leaf.double().mean()), float(leaf.max()))
            else:
                good = leaf[fin]
                stats = (float(good.min()), float(good.double().mean()), float(good.max()))
        else:
            a = np.asarray(leaf)
            n_bad = int(np.sum(~np.isfinite(a)))
            stats = ((float("nan"),) * 3 if n_bad == a.size else
                     (float(np.nanmin(a)), float(np.nanmean(a)), float(np.nanmax(a))))
        out[f.name] = {"min": stats[0], "mean": stats[1], "max": stats[2], "nan": n_bad}
    return out


def _format_state_stats(diag) -> str:
    return " ".join(format_field_stats(k, min_val=s["min"], mean_val=s["mean"], max_val=s["max"],
                                       nan_count=s["nan"]) for k, s in diag.items())


def _physical_scalars(model, state) -> dict[str, float]:
    """Energy / enstrophy scalars for the run-log line from the fused device reduction."""
    try:
        a, b, _bad = model.diag_scalars(state)
    except Exception as exc:      # a corrupted state must not take the log line down
        return {"_error": f"diagnose failed: {type(exc).__name__}: {exc}"}
    a, b = np.atleast_1d(a), np.atleast_1d(b)
    first = "energy" if hasattr(state, "h") else "kinetic_energy"
    flat = {}
    if a.size == 1:
        flat[first], flat["enstrophy"] = float(a[0]), float(b[0])
    else:
        for i in range(a.size):
            flat[f"{first}_layer_{i}"] = float(a[i])
            flat[f"enstrophy_layer_{i}"] = float(b[i])
        flat[f"total_{first}"], flat["total_enstrophy"] = float(a.sum()), float(b.sum())
    return flat


_HEADLINE = ("total_energy", "total_kinetic_energy", "energy", "kinetic_energy", "total_enstrophy", "enstrophy")


def _format_physical(flat) -> str:
    if "_error" in flat:
        return flat["_error"]
    return " ".join(f"{k}={flat[k]:.3g}" for k in _HEADLINE if k in flat)


def _to_device_state(state, dtype):
    """Keep the state on the device between chunks."""
    import torch
    kw = {}
    for f in dataclasses.fields(state):
        x = getattr(state, f.name)
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
        kw[f.name] = t.to("cuda", dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.float64)
    return type(state)(**kw)


def _attrs_for(spec: RunSpec, *, mode: str) -> dict[str, Any]:
    ts = spec.timestepping
    return {"somax_sim_mode": mode, "testcase_name": spec.testcase.name, "t0": float(ts.t0),
            "t1": float(ts.t1), "dt": float(ts.dt), "save_interval": float(ts.save_interval)}


def _chunked_integrate_with_diagnostics(model, state0, save_ts, dt, *, diagnostics_per_save,
                                        max_steps_per_chunk, run_log, mode, on_snapshot=None):
    """Integrate chunk by chunk over the diagnostics grid (cli/_run.py:540-700).  `on_snapshot(i,
    t, state)` is called for every save time (including t0) as soon as its state exists.  Returns
    (final_state, n_steps)."""
    diag_ts, save_idx = _build_diagnostic_grid(np.asarray(save_ts), diagnostics_per_save)
    n_int = diag_ts.shape[0] - 1
    sd = _state_diagnostics(state0)
    phys = _format_physical(_physical_scalars(model, state0))
    run_log.debug(f"chunk 0/{n_int} sim_t={format_time_seconds(float(diag_ts[0]))} | {_format_state_stats(sd)}"
                  + (f" | physics: {phys}" if phys else "") + " | initial state")
    if on_snapshot:
        on_snapshot(0, float(diag_ts[0]), state0)
    state, prev_energy, n_steps = state0, None, 0
    t_all = time.perf_counter()
    for i in range(n_int):
        c0, c1 = float(diag_ts[i]), float(diag_ts[i + 1])
        t_c = time.perf_counter()
        sol = model.integrate(state, c0, c1, dt, max_steps=max_steps_per_chunk)
        new_state = type(state0)(**{f.name: getattr(sol.ys, f.name)[-1] for f in dataclasses.fields(state0)})
        sd = _state_diagnostics(new_state)           # synchronises: the chunk has finished here
        wall = time.perf_counter() - t_c
        n_steps += int(sol.stats.get("num_steps", 0))
        flat = _physical_scalars(model, new_state)
        energy = next((flat[k] for k in ("total_energy", "energy", "total_kinetic_energy", "kinetic_energy")
                       if k in flat), None)
        warn = ""
        if prev_energy and energy is not None and not np.isnan(energy) and abs(energy) > 10 * abs(prev_energy):
            warn = f" !! energy grew {energy / prev_energy:.1f}x from previous chunk"
        if energy is not None and not np.isnan(energy):
            prev_energy = energy
        phys = _format_physical(flat)
        run_log.debug(f"chunk {i + 1}/{n_int} sim_t={format_time_seconds(c1)} | {_format_state_stats(sd)}"
                      + (f" | physics: {phys}" if phys else "") + f" | wall={format_wallclock(wall)}" + warn)
        if any(s["nan"] > 0 for s in sd.values()):
            run_log.debug(f"ABORT at chunk {i + 1}/{n_int}: non-finite state")
            raise IntegrationDivergedError(
                f"somax-sim {mode} integration produced non-finite values during chunk {i + 1}/{n_int} "
                f"(sim_t={format_time_seconds(c1)}).\n  {_format_state_stats(sd)}\n"
                f"  Refusing to write artifacts. The non-finite values appeared between "
                f"sim_t={format_time_seconds(c0)} and sim_t={format_time_seconds(c1)} — earlier chunks were "
                f"finite. This is consistent with a slow numerical instability, not an immediate CFL violation.")
        if (i + 1) in save_idx and on_snapshot:
            on_snapshot(sorted(save_idx).index(i + 1), c1, new_state)
        state = new_state
    run_log.debug(f"all {n_int} chunks completed in {format_wallclock(time.perf_counter() - t_all)}")
    return state, n_steps


def _assert_finite_state(stacked, *, mode: str) -> None:
    """Unconditional safety net on a time-stacked state (cli/_run.py:703-757): raise
    `IntegrationDivergedError` naming the fields with non-finite values, and say whether the first
    saved slice was already bad (CFL violation at t0 / bad initial state) or the run blew up later."""
    bad, first_bad = [], []
    for f in dataclasses.fields(stacked):
        leaf = getattr(stacked, f.name)
        if hasattr(leaf, "is_cuda"):
            import torch
            nonfin = ~torch.isfinite(leaf)
            n_bad, size = int(nonfin.sum()), leaf.numel()
            n_first, size_first = (int(nonfin[0].sum()), leaf[0].numel()) if leaf.shape[0] >= 1 else (0, 0)
        else:
            a = np.asarray(leaf)
            nonfin = ~np.isfinite(a)
            n_bad, size = int(nonfin.sum()), a.size
            n_first, size_first = (int(nonfin[0].sum()), a[0].size) if a.shape[0] >= 1 else (0, 0)
        if n_bad:
            bad.append(f"{f.name} ({n_bad}/{size} non-finite)")
            if n_first:
                first_bad.append(f"{f.name} ({n_first}/{size_first} at t0)")
    if not bad:
        return
    lines = [f"somax-sim {mode} integration produced non-finite values:", "  " + "; ".join(bad)]
    if first_bad:
        lines += ["  → already non-finite at the first saved step:", "    " + "; ".join(first_bad),
                  "  This usually means CFL violation at t0 or bad initial conditions."]
    else:
        lines.append("  Mid-integration blow-up. Most common cause: CFL violation "
                     "(time step too large for the grid spacing and the fastest wave).")
    lines.append("  Refusing to write artifacts. Check the timestepping (dt vs dx, wave speeds) and try a smaller dt.")
    raise IntegrationDivergedError("\n".join(lines))


def _integrate_and_write(spec: RunSpec, output_dir: Path, *, mode: str, initial_state,
                         diagnostics_per_save: int = 1) -> SimulationResult:
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    adapter = get_adapter(spec.testcase.name)
    model, factory_state0 = adapter(grid=spec.testcase.grid, consts=spec.testcase.consts,
                                    stratification=spec.testcase.stratification, params=spec.testcase.params)
    state0 = initial_state if initial_state is not None else factory_state0
    _assertions.run_preflight(spec, model)      # CFL etc. before any compute is spent
    state0 = _to_device_state(state0, model.dtype)
    only_final = mode == "spinup"
    save_ts = _build_save_times(spec, only_final=only_final)
    ts = spec.timestepping
    expected = int((ts.t1 - ts.t0) / ts.dt)
    max_steps = max(16384, int(expected * 1.2))
    n_diag = max(len(save_ts) - 1, 1) * max(int(diagnostics_per_save), 1)
    per_chunk_max = max(2048, max_steps // n_diag + 256)
    run_log = RunLog(output_dir, label=f"somax-sim/{mode}")
    run_log.debug(f"integration starting: testcase={spec.testcase.name} t0={format_time_seconds(ts.t0)} "
                  f"t1={format_time_seconds(ts.t1)} dt={ts.dt} s save_n={len(save_ts)} "
                  f"diagnostics_per_save={diagnostics_per_save}")
    # snapshots stream into a scratch store while the run is in flight; it is renamed into place
    # only after the finite-state check and the run completes (no artifacts on failure)
    write_snaps = spec.output.write_snapshots and not only_final
    tmp_snap = output_dir / ".snapshots.zarr.partial"
    writer = io.AsyncSnapshotWriter(tmp_snap, type(state0), attrs=_attrs_for(spec, mode=mode)) if write_snaps else None
    t_start = time.perf_counter()
    try:
        final_state, n_steps = _chunked_integrate_with_diagnostics(
            model, state0, save_ts, ts.dt, diagnostics_per_save=diagnostics_per_save,
            max_steps_per_chunk=per_chunk_max, run_log=run_log, mode=mode,
            on_snapshot=(lambda i, t, st: writer.put(t, st)) if writer else None)
    except Exception as exc:
        if writer:
            try:
                writer.close()
            except Exception:
                pass
            _rmtree(tmp_snap)
        run_log.stop(f"FAILED during integrate: {type(exc).__name__}: {exc}")
        raise
    wallclock = time.perf_counter() - t_start
    run_log.debug(f"integration finished in {format_wallclock(wallclock)}")
    try:
        if writer:
            writer.close()
        _assert_finite_state(type(final_state)(**{f.name: getattr(final_state, f.name)[None]
                                                  for f in dataclasses.fields(final_state)}), mode=mode)
        metrics: dict[str, Any] = {}
        if spec.output.write_metrics and not only_final:
            try:
                metrics = _flatten_diagnostics(model.diagnose(final_state))
            except Exception as exc:
                logger.warning("model.diagnose failed: %s", exc)
            metrics.update(wallclock_seconds=wallclock, n_steps=n_steps, t0=ts.t0, t1=ts.t1,
                           save_interval=ts.save_interval, mode=mode)
        _assertions.run_postflight(spec, metrics)   # before any artifact is moved into place
        snapshots_path = None
        if write_snaps:
            snapshots_path = output_dir / "snapshots.zarr"
            _rmtree(snapshots_path)
            tmp_snap.rename(snapshots_path)
            run_log.debug(f"wrote {snapshots_path.name}")
        final_state_path = output_dir / "final_state.zarr"
        io.save_dataset(io.state_to_dataset(final_state, time=float(ts.t1), attrs=_attrs_for(spec, mode=mode)),
                        final_state_path, mode="w")
        run_log.debug(f"wrote {final_state_path.name}")
        metrics_path = None
        if spec.output.write_metrics and not only_final and metrics:
            metrics_path = output_dir / "metrics.json"
            metrics_path.write_text(json.dumps(metrics, indent=2, sort_keys=True))
        dump_yaml(spec, str(output_dir / "resolved.yaml"))
    except Exception as exc:
        _rmtree(tmp_snap)       # no artifacts on failure, the scratch snapshot store included
        run_log.stop(f"FAILED postflight: {type(exc).__name__}: {exc}")
        raise
    run_log.stop("finished cleanly")
    return SimulationResult(output_dir, snapshots_path, final_state_path, metrics_path, wallclock, n_steps)


def _rmtree(p: Path):
    import shutil
    if Path(p).exists():
        shutil.rmtree(p)


def simulate(spec: RunSpec, output_dir, *, diagnostics_per_save: int = 1) -> SimulationResult:
    """Fresh run from the factory initial state (cli/_run.py:783-807)."""
    return _integrate_and_write(spec, Path(output_dir), mode="run", initial_state=None,
                                diagnostics_per_save=diagnostics_per_save)


def spinup(spec: RunSpec, output_dir, *, diagnostics_per_save: int = 1) -> SimulationResult:
    """Spinup: only `final_state.zarr` (the restart artifact) is written (cli/_run.py:810-832)."""
    return _integrate_and_write(spec, Path(output_dir), mode="spinup", initial_state=None,
                                diagnostics_per_save=diagnostics_per_save)


def restart(spec: RunSpec, output_dir, *, restart_from, diagnostics_per_save: int = 1) -> SimulationResult:
    """Continue from a saved state (cli/_run.py:835-878); the store must hold the test case's
    state class."""
    ds = io.load_dataset(Path(restart_from))
    adapter = get_adapter(spec.testcase.name)
    _m, factory_state0 = adapter(grid=spec.testcase.grid, consts=spec.testcase.consts,
                                 stratification=spec.testcase.stratification, params=spec.testcase.params)
    expected = type(factory_state0)
    stored = ds.attrs.get("state_class")
    if stored is not None and stored != expected.__name__:
        raise TypeError(f"restart artifact at {restart_from} contains a {stored}, but the testcase "
                        f"{spec.testcase.name!r} expects {expected.__name__}. Check that --from points at a "
                        "final_state.zarr written by a compatible run.")
    state0 = io.dataset_to_state(ds, state_class=expected)
    return _integrate_and_write(spec, Path(output_dir), mode="restart", initial_state=state0,
                                diagnostics_per_save=diagnostics_per_save)
