"""`somax-sim` command line on the B200 path: `python -m somax_b200.cli.app <command> ...`.

argparse stand-in for the reference's cyclopts app (somax/_src/cli/app.py): same commands (run,
spinup, restart, list-testcases, list-models, show-config) and options (--config, --output-dir,
--debug, --diagnostics-per-save, --verbose, --from)."""
from __future__ import annotations

import argparse
import logging
import sys
from pathlib import Path

from . import _factories, _run
from .spec import RunSpec, load_yaml


def _load_and_prepare(config: Path, *, debug: bool) -> RunSpec:
    spec = load_yaml(str(config))
    if debug:
        spec = spec.with_debug_applied()
        spec.validate()
    return spec


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="somax-sim", description="somax simulation runner — fresh runs, spinups, restarts.")
    sub = ap.add_subparsers(dest="command", required=True)
    for name in ("run", "spinup", "restart"):
        p = sub.add_parser(name)
        p.add_argument("--config", type=Path, required=True, help="Path to a YAML run-spec file.")
        p.add_argument("--output-dir", type=Path, required=True,
                       help="Directory for snapshots.zarr / final_state.zarr / metrics.json.")
        p.add_argument("--debug", action="store_true", help="Apply the cfg.debug overrides (smaller grid, shorter run).")
        p.add_argument("--diagnostics-per-save", type=int, default=1,
                       help="Diagnostic sub-chunks logged per save interval in <output_dir>/run.log.")
        p.add_argument("--verbose", action="store_true", help="DEBUG-level logging on stderr (tees run.log lines).")
        if name == "restart":
            p.add_argument("--from", dest="from_", type=Path, required=True,
                           help="zarr store holding the state to restart from (a final_state.zarr).")
    sub.add_parser("list-testcases")
    sub.add_parser("list-models")
    sc = sub.add_parser("show-config")
    sc.add_argument("path", type=Path)
    return ap


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.command == "list-testcases":
        print("Registered test cases:")
        for n in _factories.list_test_cases():
            print(f"  - {n}")
        return 0
    if args.command == "list-models":
        import somax_b200 as sb
        print("Available model classes:")
        for n in sorted(n for n in dir(sb) if isinstance(getattr(sb, n), type) and issubclass(getattr(sb, n), sb.SomaxModel)
                        and getattr(sb, n) is not sb.SomaxModel):
            print(f"  - {n}")
        return 0
    if args.command == "show-config":
        import yaml
        spec = load_yaml(str(args.path))
        print(f"# Resolved RunSpec from {args.path}")
        print(yaml.safe_dump(spec.to_dict(), sort_keys=False, default_flow_style=False))
        return 0
    logging.basicConfig(stream=sys.stderr, level=logging.DEBUG if args.verbose else logging.INFO,
                        format="%(asctime)s | %(levelname)-7s | %(message)s")
    spec = _load_and_prepare(args.config, debug=args.debug)
    kw = dict(diagnostics_per_save=args.diagnostics_per_save)
    if args.command == "run":
        _run.simulate(spec, args.output_dir, **kw)
    elif args.command == "spinup":
        _run.spinup(spec, args.output_dir, **kw)
    else:
        _run.restart(spec, args.output_dir, restart_from=args.from_, **kw)
    return 0


if __name__ == "__main__":   # pragma: no cover
    sys.exit(main())
