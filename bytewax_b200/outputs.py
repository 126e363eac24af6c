"""Output sinks: host-side mirror of ``bytewax.outputs`` (pysrc/bytewax/outputs.py)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Generic, List, Optional, TypeVar
from zlib import adler32

X = TypeVar("X")
S = TypeVar("S")


class Sink(ABC, Generic[X]):  # noqa: B024
    """A destination to write items to."""


class StatefulSinkPartition(ABC, Generic[X, S]):
    @abstractmethod
    def write_batch(self, values: List[X]) -> None: ...

    @abstractmethod
    def snapshot(self) -> S: ...

    def close(self) -> None:
        return


class FixedPartitionedSink(Sink, Generic[X, S]):
    @abstractmethod
    def list_parts(self) -> List[str]: ...

    def part_fn(self, item_key: str) -> int:
        """Route a key to a partition; stable across processes (outputs.py:100-127)."""
        return adler32(item_key.encode())

    @abstractmethod
    def build_part(self, step_id: str, for_part: str, resume_state: Optional[S]) -> StatefulSinkPartition[X, S]: ...


class StatelessSinkPartition(ABC, Generic[X]):
    @abstractmethod
    def write_batch(self, items: List[X]) -> None: ...

    def close(self) -> None:
        return


class DynamicSink(Sink[X]):
    @abstractmethod
    def build(self, step_id: str, worker_index: int, worker_count: int) -> StatelessSinkPartition[X]: ...
