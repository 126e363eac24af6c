"""``bytewax.recovery`` (pysrc/bytewax/recovery.py:6-19): the names load; the store itself is out of scope."""
from bytewax_b200._bytewax import (  # noqa: F401
    InconsistentPartitionsError,
    MissingPartitionsError,
    NoPartitionsError,
    RecoveryConfig,
    init_db_dir,
)

__all__ = ["InconsistentPartitionsError", "NoPartitionsError", "MissingPartitionsError", "RecoveryConfig", "init_db_dir"]
