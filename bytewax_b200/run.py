"""``python -m bytewax_b200.run <module[:attr_or_call]>``: mirror of ``bytewax.run``.

Locates a ``Dataflow`` the way pysrc/bytewax/run.py:30-117 does (module path or
dotted name, optional ``:variable`` or ``:factory(args)``) and hands it to the
host engine (``bytewax_b200.engine.cli_main``).  Recovery and multi-process
flags are parsed for compatibility; using them reports that they are out of scope.
"""

from __future__ import annotations

import argparse
import ast
import importlib
import inspect
import os
import sys
from datetime import timedelta
from pathlib import Path
from typing import Optional

from bytewax_b200.dataflow import Dataflow
from bytewax_b200.engine import cli_main

__all__ = ["cli_main"]


class _EnvDefault(argparse.Action):
    """Take the default of an option from an environment variable (run.py:140-150)."""

    def __init__(self, envvar, default=None, **kwargs):
        if envvar:
            default = os.environ.get(envvar, default)
            kwargs["help"] = f"{kwargs.get('help', '')} [env: {envvar}]"
        super().__init__(default=default, **kwargs)

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, values)


def _locate_dataflow(module_name: str, dataflow_name: str):
    """Import ``module_name`` and resolve ``dataflow_name`` (a variable or a call with literal args)."""
    try:
        __import__(module_name)
    except ImportError as ex:
        if ex.__traceback__ is not None and ex.__traceback__.tb_next is not None:
            raise
        raise ImportError(f"Could not import {module_name!r}.") from None
    module = sys.modules[module_name]
    try:
        expr = ast.parse(dataflow_name.strip(), mode="eval").body
    except SyntaxError:
        raise SyntaxError(f"Failed to parse {dataflow_name!r} as an attribute name or function call") from None
    if isinstance(expr, ast.Name):
        name, args, kwargs = expr.id, [], {}
    elif isinstance(expr, ast.Call):
        if not isinstance(expr.func, ast.Name):
            raise TypeError(f"Function reference must be a simple name: {dataflow_name!r}")
        name = expr.func.id
        try:
            args = [ast.literal_eval(a) for a in expr.args]
            kwargs = {k.arg: ast.literal_eval(k.value) for k in expr.keywords}
        except ValueError:
            raise ValueError(f"Failed to parse arguments as literal values: {dataflow_name!r}") from None
    else:
        raise ValueError(f"Failed to parse {dataflow_name!r} as an attribute name or function call")
    try:
        attr = getattr(module, name)
    except AttributeError as ex:
        raise AttributeError(f"Failed to find attribute {name!r} in {module.__name__!r}.") from ex
    if inspect.isfunction(attr):
        try:
            flow = attr(*args, **kwargs)
        except TypeError as ex:
            if not _called_with_wrong_args(attr):
                raise
            raise TypeError(f"The factory {dataflow_name!r} in module {module.__name__!r} could not be called with the specified arguments") from ex
    else:
        flow = attr
    if isinstance(flow, Dataflow):
        return flow
    raise RuntimeError("A valid Bytewax dataflow was not obtained from " f"'{module.__name__}:{dataflow_name}'")


def _called_with_wrong_args(f) -> bool:
    tb = sys.exc_info()[2]
    try:
        while tb is not None:
            if tb.tb_frame.f_code is f.__code__:
                return False
            tb = tb.tb_next
        return True
    finally:
        del tb


def _prepare_import(import_str: str):
    """``path/to/flow.py:attr`` or ``pkg.mod:attr`` -> (module name, attribute) (run.py:153-190)."""
    path, _, flow_name = import_str.partition(":")
    if flow_name == "":
        flow_name = "flow"
    path = os.path.realpath(path) if (os.path.sep in path or path.endswith(".py")) else path
    if os.path.sep in path or path.endswith(".py"):
        fname, ext = os.path.splitext(path)
        if ext == ".py":
            path = fname
        if os.path.basename(path) == "__init__":
            path = os.path.dirname(path)
        parts = []
        while True:
            path, name = os.path.split(path)
            parts.append(name)
            if not os.path.exists(os.path.join(path, "__init__.py")):
                break
        if sys.path[0] != path:
            sys.path.insert(0, path)
        return ".".join(parts[::-1]) + ":" + flow_name
    if sys.path[0] != os.getcwd():
        sys.path.insert(0, os.getcwd())
    return path + ":" + flow_name


def _parse_timedelta(s: str) -> timedelta:
    return timedelta(seconds=float(s))


def _create_arg_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="python -m bytewax_b200.run", description="Run a dataflow")
    p.add_argument("import_str", type=str, help="Dataflow import string in the format <module_name>[:<dataflow_variable_or_factory>]")
    scaling = p.add_argument_group("Scaling")
    scaling.add_argument("-w", "--workers-per-process", type=int, action=_EnvDefault, envvar="BYTEWAX_WORKERS_PER_PROCESS",
                         help="Number of workers for each process")
    scaling.add_argument("-i", "--process-id", type=int, action=_EnvDefault, envvar="BYTEWAX_PROCESS_ID", help="Process id")
    scaling.add_argument("-a", "--addresses", action=_EnvDefault, envvar="BYTEWAX_ADDRESSES",
                         help="Addresses of other processes, separated by semicolon")
    rec = p.add_argument_group("Recovery")
    rec.add_argument("-r", "--recovery-directory", type=Path, action=_EnvDefault, envvar="BYTEWAX_RECOVERY_DIRECTORY",
                     help="Local file system directory to look for pre-initialized recovery partitions")
    rec.add_argument("-s", "--snapshot-interval", type=_parse_timedelta, action=_EnvDefault, envvar="BYTEWAX_SNAPSHOT_INTERVAL",
                     help="System time duration in seconds to snapshot state for recovery")
    rec.add_argument("-b", "--backup-interval", type=_parse_timedelta, action=_EnvDefault, envvar="BYTEWAX_RECOVERY_BACKUP_INTERVAL",
                     help="System time duration in seconds to keep extra state snapshots around")
    return p


def _parse_args(argv=None):
    args = _create_arg_parser().parse_args(argv)
    if args.workers_per_process is not None:
        args.workers_per_process = int(args.workers_per_process)
    if args.process_id is not None:
        args.process_id = int(args.process_id)
    if args.recovery_directory is not None and args.snapshot_interval is None:
        _create_arg_parser().error("when running with recovery, the `-s/--snapshot-interval` value must be set")
    return args


def main(argv=None):
    args = _parse_args(argv)
    if args.recovery_directory is not None:
        raise NotImplementedError("recovery is out of scope of this engine (SURVEY.md section 2 row 12)")
    mod_str, _, attr_str = _prepare_import(args.import_str).partition(":")
    flow = _locate_dataflow(mod_str, attr_str)
    addresses = args.addresses.split(";") if args.addresses else None
    cli_main(flow, workers_per_process=args.workers_per_process or 1, process_id=args.process_id, addresses=addresses,
             epoch_interval=args.snapshot_interval)


if __name__ == "__main__":
    main()
