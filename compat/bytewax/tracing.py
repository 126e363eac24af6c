"""Alias of ``bytewax_b200.tracing``: with ``compat/`` on PYTHONPATH, flows written against ``bytewax`` load unchanged."""
import sys as _sys

import bytewax_b200.tracing as _impl

_sys.modules[__name__] = _impl
