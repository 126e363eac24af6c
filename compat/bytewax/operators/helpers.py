"""Alias of ``bytewax_b200.operators.helpers``: with ``compat/`` on PYTHONPATH, flows written against ``bytewax`` load unchanged."""
import sys as _sys

if __name__ == "__main__":
    import runpy as _runpy

    _runpy.run_module("bytewax_b200.operators.helpers", run_name="__main__", alter_sys=True)
else:
    import bytewax_b200.operators.helpers as _impl

    _sys.modules[__name__] = _impl
