"""Opt-in preflight / postflight assertions of a run (mirror of somax/_src/cli/_assertions.py):
`cfl` before any stepping, `bounded_metric` on the metrics dict before it is written.  Failures
raise `AssertionFailedError`; unknown names are failures too (config typos)."""
from __future__ import annotations

import math
from typing import Any, Callable


class AssertionFailedError(RuntimeError):
    """An opt-in preflight or postflight assertion failed."""


def check_cfl(spec, model, *, wave_speed_m_per_s: float, max_cfl: float = 0.5) -> None:
    """CFL = wave_speed * dt / min(dx, dy) must not exceed max_cfl (_assertions.py:49-103)."""
    if wave_speed_m_per_s <= 0:
        raise AssertionFailedError(f"cfl: wave_speed_m_per_s must be > 0 (got {wave_speed_m_per_s})")
    if max_cfl <= 0:
        raise AssertionFailedError(f"cfl: max_cfl must be > 0 (got {max_cfl})")
    grid = getattr(model, "grid", None)
    if grid is None:
        raise AssertionFailedError(f"cfl: model {type(model).__name__!r} has no .grid attribute; cannot infer dx")
    dx_min, dt = float(min(grid.dx, grid.dy)), float(spec.timestepping.dt)
    cfl = wave_speed_m_per_s * dt / dx_min
    if cfl > max_cfl:
        raise AssertionFailedError(
            f"cfl check FAILED: CFL = {cfl:.3f} > max_cfl = {max_cfl}\n"
            f"  wave_speed = {wave_speed_m_per_s:.2f} m/s\n  dt         = {dt:.4f} s\n"
            f"  dx_min     = {dx_min:.2f} m\n"
            f"  → maximum stable dt at this CFL: {max_cfl * dx_min / wave_speed_m_per_s:.4f} s")


def check_bounded_metric(spec, metrics: dict, *, name: str, min: float | None = None,
                         max: float | None = None) -> None:
    """metrics[name] must be a finite scalar inside [min, max] (_assertions.py:116-160)."""
    if name not in metrics:
        raise AssertionFailedError(f"bounded_metric: metric {name!r} not present in run output. "
                                   f"Available metrics: {sorted(metrics)}")
    raw = metrics[name]
    try:
        value = float(raw)
    except (TypeError, ValueError) as exc:
        raise AssertionFailedError(f"bounded_metric: metric {name!r} is not a numeric scalar (got {raw!r})") from exc
    if not math.isfinite(value):
        raise AssertionFailedError(f"bounded_metric: metric {name!r} is non-finite ({value})")
    if min is not None and value < min:
        raise AssertionFailedError(f"bounded_metric: {name} = {value} is below min = {min}")
    if max is not None and value > max:
        raise AssertionFailedError(f"bounded_metric: {name} = {value} is above max = {max}")


PREFLIGHT_ASSERTIONS: dict[str, Callable[..., None]] = {"cfl": check_cfl}
POSTFLIGHT_ASSERTIONS: dict[str, Callable[..., None]] = {"bounded_metric": check_bounded_metric}


def _run(phase: dict, other: dict, spec, subject: Any) -> None:
    for name, params in (spec.assertions or {}).items():
        if name in other:
            continue
        check = phase.get(name)
        if check is None:
            raise AssertionFailedError(
                f"unknown assertion {name!r}; available preflight: {sorted(PREFLIGHT_ASSERTIONS)}; "
                f"available postflight: {sorted(POSTFLIGHT_ASSERTIONS)}")
        check(spec, subject, **(params or {}))


def run_preflight(spec, model) -> None:
    _run(PREFLIGHT_ASSERTIONS, POSTFLIGHT_ASSERTIONS, spec, model)


def run_postflight(spec, metrics: dict) -> None:
    _run(POSTFLIGHT_ASSERTIONS, PREFLIGHT_ASSERTIONS, spec, metrics)
