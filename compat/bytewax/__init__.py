"""Alias of ``bytewax_b200``: with ``compat/`` on PYTHONPATH, flows written against ``bytewax`` load unchanged.

The package object stays this stub (so that the sub-module aliases next to it are found); every name of the
implementation, private ones included, is re-exported."""
import bytewax_b200 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not (k.startswith("__") and k.endswith("__"))})
