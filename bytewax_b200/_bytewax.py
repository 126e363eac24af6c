"""Names of the reference's PyO3 module ``bytewax._bytewax`` (src/lib.rs:24-32) that flows import through
``bytewax.recovery`` / ``bytewax.tracing`` / ``bytewax.run``: enough for a flow that merely imports or constructs them
to load.  The subsystems behind them (SQLite recovery store, tracing exporters) are out of scope:
a ``RecoveryConfig`` handed to ``run_main`` is refused there, loudly."""
from datetime import timedelta
from pathlib import Path
from typing import Optional

from bytewax_b200.engine import cli_main, cluster_main, run_main  # noqa: F401
from bytewax_b200.inputs import AbortExecution  # noqa: F401


class InconsistentPartitionsError(ValueError):
    """(src/recovery.rs) Two recovery partitions disagree about the state of the cluster."""


class MissingPartitionsError(FileNotFoundError):
    """(src/recovery.rs) Not all recovery partitions can be found."""


class NoPartitionsError(FileNotFoundError):
    """(src/recovery.rs) No recovery partitions at all in ``db_dir``."""


class RecoveryConfig:
    """``RecoveryConfig(db_dir, backup_interval=None)`` (src/recovery.rs:123-160): holds its arguments."""

    def __init__(self, db_dir, backup_interval: Optional[timedelta] = None):
        self.db_dir = Path(db_dir)
        self.backup_interval = backup_interval if backup_interval is not None else timedelta(0)


def init_db_dir(db_dir, count: int) -> None:
    raise NotImplementedError("the SQLite recovery store is out of scope of this engine (SURVEY.md section 2 row 12)")


class TracingConfig:
    """Base of the tracing configurations (src/tracing.rs): holds nothing; ``setup_tracing`` is a no-op."""


class JaegerConfig(TracingConfig):
    def __init__(self, service_name: str, endpoint: Optional[str] = None, sampling_ratio: float = 1.0):
        self.service_name, self.endpoint, self.sampling_ratio = service_name, endpoint, sampling_ratio


class OtlpTracingConfig(TracingConfig):
    def __init__(self, service_name: str, url: Optional[str] = None, sampling_ratio: float = 1.0):
        self.service_name, self.url, self.sampling_ratio = service_name, url, sampling_ratio


class BytewaxTracer:
    """Handle returned by ``setup_tracing``; keeps nothing alive here."""


def setup_tracing(tracing_config: Optional[TracingConfig] = None, log_level: Optional[str] = None) -> BytewaxTracer:
    return BytewaxTracer()
