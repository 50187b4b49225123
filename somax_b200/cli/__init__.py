"""somax-sim run layer on the B200 path: RunSpec, test-case registry, chunked runner.

Host-side mirror of somax/_src/cli/{spec.py,_factories.py,_run.py}: same dataclasses, registry
names, artifact layout (snapshots.zarr, final_state.zarr, metrics.json, resolved.yaml, run.log)
and failure semantics (IntegrationDivergedError).  The stepping itself is `model.integrate`
(libsomax_b200); between chunks only a handful of device-reduced scalars reach the host.
"""
from .spec import (DebugSpec, OutputSpec, RunSpec, TestCaseSpec, TimesteppingSpec, dump_yaml,  # noqa: F401
                   load_yaml)
from ._factories import TEST_CASES, get_adapter, list_test_cases  # noqa: F401
from ._run import (IntegrationDivergedError, SimulationResult, restart, simulate, spinup)  # noqa: F401
from ._assertions import AssertionFailedError  # noqa: F401,E402
