"""``bytewax.tracing`` (pysrc/bytewax/tracing.py:3-15): the names load; exporters are out of scope."""
from bytewax_b200._bytewax import JaegerConfig, OtlpTracingConfig, TracingConfig, setup_tracing  # noqa: F401

__all__ = ["TracingConfig", "JaegerConfig", "OtlpTracingConfig", "setup_tracing"]
