"""Alias of ``bytewax_b200._bytewax``: with ``compat/`` on PYTHONPATH, flows written against ``bytewax`` load unchanged."""
import sys as _sys

import bytewax_b200._bytewax as _impl

_sys.modules[__name__] = _impl
