"""Engine errors (mirror of pysrc/bytewax/errors.py)."""


class BytewaxRuntimeError(RuntimeError):
    """Raised by the engine; the user's exception is the ``__cause__`` (src/errors.rs:69-105)."""
