"""Run specification of one simulation (mirror of somax/_src/cli/spec.py:22-339).

A `RunSpec` is four blocks: which test case and its structured kwargs (`grid` / `consts` /
`stratification` / `params`), the integration window and snapshot cadence, output toggles, and
optional `debug` overrides that `with_debug_applied()` merges on top.  YAML round trip through
`load_yaml` / `dump_yaml`; `from_dict` ignores unknown keys, `validate` raises `ValueError`.
"""
from __future__ import annotations

import copy
from dataclasses import asdict, dataclass, field
from typing import Any

_BLOCKS = ("grid", "consts", "stratification", "params")


@dataclass
class TestCaseSpec:
    """Registry key + the four kwargs blocks handed to the factory adapter (spec.py:22-43)."""
    __test__ = False     # not a pytest class

    name: str
    grid: dict = field(default_factory=dict)
    consts: dict = field(default_factory=dict)
    stratification: dict = field(default_factory=dict)
    params: dict = field(default_factory=dict)


@dataclass
class TimesteppingSpec:
    """Seconds; `save_interval` is the snapshot spacing (spec.py:46-66)."""
    t0: float
    t1: float
    dt: float
    save_interval: float


@dataclass
class OutputSpec:
    write_snapshots: bool = True
    write_metrics: bool = True


@dataclass
class DebugSpec:
    """Overrides applied by `--debug`: per-block dict merge for `testcase`, per-key for
    `timestepping` (spec.py:86-111)."""
    testcase: dict = field(default_factory=dict)
    timestepping: dict = field(default_factory=dict)


@dataclass
class RunSpec:
    testcase: TestCaseSpec
    timestepping: TimesteppingSpec
    output: OutputSpec = field(default_factory=OutputSpec)
    debug: DebugSpec = field(default_factory=DebugSpec)
    assertions: dict = field(default_factory=dict)

    def validate(self) -> None:
        """Window and cadence sanity (spec.py:146-177); factory kwargs are checked by the factory."""
        ts = self.timestepping
        if ts.t1 <= ts.t0:
            raise ValueError(f"timestepping.t1 ({ts.t1}) must be > timestepping.t0 ({ts.t0})")
        if ts.dt <= 0:
            raise ValueError(f"timestepping.dt ({ts.dt}) must be > 0")
        if ts.save_interval <= 0:
            raise ValueError(f"timestepping.save_interval ({ts.save_interval}) must be > 0")
        if ts.save_interval > ts.t1 - ts.t0:
            raise ValueError(f"timestepping.save_interval ({ts.save_interval}) cannot exceed "
                             f"the integration window ({ts.t1 - ts.t0})")
        if not isinstance(self.testcase.name, str) or not self.testcase.name:
            raise ValueError("testcase.name must be a non-empty string")

    def with_debug_applied(self) -> "RunSpec":
        """New spec with the debug overrides merged in and consumed; `self` when there are none
        (spec.py:183-238)."""
        if not self.debug.testcase and not self.debug.timestepping:
            return self
        tc = copy.deepcopy(self.testcase)
        for block, override in self.debug.testcase.items():
            if not hasattr(tc, block):
                raise ValueError(f"debug.testcase.{block!r} does not match any TestCaseSpec field")
            target = getattr(tc, block)
            if not isinstance(target, dict) or not isinstance(override, dict):
                raise ValueError(f"debug.testcase.{block!r} merge requires both sides to be dicts; got "
                                 f"{type(target).__name__} and {type(override).__name__}")
            target.update(override)
        ts = TimesteppingSpec(**{k: self.debug.timestepping.get(k, v)
                                 for k, v in asdict(self.timestepping).items()})
        return RunSpec(testcase=tc, timestepping=ts, output=self.output, debug=DebugSpec())

    def to_dict(self) -> dict[str, Any]:
        d = {"testcase": {"name": self.testcase.name}}
        for b in _BLOCKS:
            d["testcase"][b] = copy.deepcopy(getattr(self.testcase, b))
        d["timestepping"] = asdict(self.timestepping)
        d["output"] = asdict(self.output)
        d["debug"] = {"testcase": copy.deepcopy(self.debug.testcase),
                      "timestepping": copy.deepcopy(self.debug.timestepping)}
        d["assertions"] = copy.deepcopy(self.assertions)
        return d

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "RunSpec":
        for required in ("testcase", "timestepping"):
            if required not in data:
                raise ValueError(f"RunSpec config missing required block: {required}")
        tc, ts = data["testcase"], data["timestepping"]
        testcase = TestCaseSpec(name=tc["name"], **{b: dict(tc.get(b, {})) for b in _BLOCKS})
        timestepping = TimesteppingSpec(*(float(ts[k]) for k in ("t0", "t1", "dt", "save_interval")))
        out, dbg = data.get("output", {}), data.get("debug", {})
        return cls(
            testcase=testcase, timestepping=timestepping,
            output=OutputSpec(bool(out.get("write_snapshots", True)), bool(out.get("write_metrics", True))),
            debug=DebugSpec(dict(dbg.get("testcase", {})), dict(dbg.get("timestepping", {}))),
            assertions={str(k): dict(v or {}) for k, v in (data.get("assertions") or {}).items()})


def load_yaml(path: str) -> RunSpec:
    import yaml
    with open(path) as f:
        data = yaml.safe_load(f)
    if not isinstance(data, dict):
        raise ValueError(f"Config file {path!r} did not parse as a top-level mapping; got {type(data).__name__}")
    spec = RunSpec.from_dict(data)
    spec.validate()
    return spec


def dump_yaml(spec: RunSpec, path: str) -> None:
    import yaml
    with open(path, "w") as f:
        yaml.safe_dump(spec.to_dict(), f, sort_keys=False, default_flow_style=False)
