"""Random metric source for demos (mirror of ``bytewax.connectors.demo``, pysrc/bytewax/connectors/demo.py).

Emits ``(metric_name, value)`` every ``interval`` for ``count`` values; one partition named after the metric, so
several metrics are several ``input`` steps merged together.  Feeds ``examples/anomaly_detector.py`` (config C2's shape).
"""

import random
import sys
from datetime import datetime, timedelta, timezone
from typing import Callable, List, Optional, Tuple

from bytewax_b200.inputs import FixedPartitionedSource, StatefulSourcePartition


class _RandomMetricPartition(StatefulSourcePartition):
    def __init__(self, name: str, interval: timedelta, count: int, next_random: Callable[[], float], resume_state):
        self._name, self._interval, self._limit, self._next_random = name, interval, count, next_random
        # resume state: (next awake time, values emitted so far)
        self._awake_at, self._emitted = resume_state if resume_state is not None else (datetime.now(timezone.utc), 0)

    def next_batch(self) -> List[Tuple[str, float]]:
        self._awake_at += self._interval
        self._emitted += 1
        if self._emitted > self._limit:
            raise StopIteration()
        return [(self._name, self._next_random())]

    def next_awake(self) -> Optional[datetime]:
        return self._awake_at

    def snapshot(self):
        return (self._awake_at, self._emitted)


class RandomMetricSource(FixedPartitionedSource):
    def __init__(self, metric_name: str, interval: timedelta = timedelta(seconds=0.7), count: int = sys.maxsize,
                 next_random: Callable[[], float] = lambda: random.randrange(0, 10)):
        self._metric_name, self._interval, self._count, self._next_random = metric_name, interval, count, next_random

    def list_parts(self) -> List[str]:
        return [self._metric_name]

    def build_part(self, step_id: str, for_part: str, resume_state):
        return _RandomMetricPartition(for_part, self._interval, self._count, self._next_random, resume_state)
