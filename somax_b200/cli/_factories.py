"""Test-case registry: structured RunSpec blocks -> `(model, state0)` (mirror of
somax/_src/cli/_factories.py:36-172; same four names, same required keys)."""
from __future__ import annotations

from typing import Any, Callable

from .. import gfd_testcases as _g

Adapter = Callable[..., tuple]

# name -> (factory, {factory kwarg: (block, key)}); tuples are made from list-valued YAML entries
_G, _C, _S, _P = "grid", "consts", "stratification", "params"
_GRID = {k: (_G, k) for k in ("nx", "ny", "Lx", "Ly")}
_ROT = {"f0": (_C, "f0"), "beta": (_C, "beta")}
_SPECS = {
    "barotropic_jet_instability": (_g.barotropic_jet_instability, {
        **_GRID, **_ROT, "H0": (_C, "H0"), "jet_speed": (_P, "jet_speed"), "jet_width": (_P, "jet_width"),
        "perturbation": (_P, "perturbation"), "lateral_viscosity": (_P, "lateral_viscosity")}),
    "doublegyre_qg": (_g.doublegyre_qg, {
        **_GRID, **_ROT, "lateral_viscosity": (_P, "lateral_viscosity"), "bottom_drag": (_P, "bottom_drag"),
        "wind_amplitude": (_P, "wind_amplitude")}),
    "doublegyre_baroclinic_qg": (_g.doublegyre_baroclinic_qg, {
        **_GRID, **_ROT, "n_layers": (_C, "n_layers"), "H": (_S, "H"), "g_prime": (_S, "g_prime"),
        "lateral_viscosity": (_P, "lateral_viscosity"), "bottom_drag": (_P, "bottom_drag"),
        "wind_amplitude": (_P, "wind_amplitude")}),
    "baroclinic_instability_swm": (_g.baroclinic_instability_swm, {
        **_GRID, **_ROT, "H": (_S, "H"), "g_prime": (_S, "g_prime"),
        "lateral_viscosity": (_P, "lateral_viscosity"), "bottom_drag": (_P, "bottom_drag"),
        "jet_speed": (_P, "jet_speed"), "jet_width": (_P, "jet_width"), "perturbation": (_P, "perturbation")}),
}


def _make_adapter(name: str) -> Adapter:
    factory, mapping = _SPECS[name]

    def adapter(*, grid: dict, consts: dict, stratification: dict, params: dict, **extra: Any):
        blocks = {_G: grid, _C: consts, _S: stratification, _P: params}
        kw = {}
        for arg, (block, key) in mapping.items():
            v = blocks[block][key]                     # KeyError names the missing entry, as in the reference
            kw[arg] = tuple(v) if isinstance(v, (list, tuple)) else v
        if "dtype" in grid:                            # extension: precision of the CUDA pipeline
            kw["dtype"] = grid["dtype"]
        return factory(**kw)

    adapter.__name__ = name
    return adapter


TEST_CASES: dict[str, Adapter] = {name: _make_adapter(name) for name in _SPECS}


def list_test_cases() -> list[str]:
    return sorted(TEST_CASES)


def get_adapter(name: str) -> Adapter:
    try:
        return TEST_CASES[name]
    except KeyError as exc:
        raise KeyError(f"unknown test case {name!r}; available: {', '.join(list_test_cases())}") from exc
