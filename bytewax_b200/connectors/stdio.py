"""Standard output sink (mirror of pysrc/bytewax/connectors/stdio.py)."""

import sys
from typing import Any, List

from bytewax_b200.outputs import DynamicSink, StatelessSinkPartition


class _PrintSinkPartition(StatelessSinkPartition[Any]):
    def write_batch(self, items: List[Any]) -> None:
        for item in items:
            sys.stdout.write(f"{item}\n")
        sys.stdout.flush()


class StdOutSink(DynamicSink[Any]):
    """Write each item's ``str`` to stdout, one per line."""

    def build(self, step_id: str, worker_index: int, worker_count: int) -> _PrintSinkPartition:
        return _PrintSinkPartition()
