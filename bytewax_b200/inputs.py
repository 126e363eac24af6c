"""Input sources: host-side mirror of ``bytewax.inputs`` (pysrc/bytewax/inputs.py).

Abstract base classes only -- the I/O itself is out of the GPU path's scope
(SURVEY.md section 2 row 10); they exist so flows written against the reference load
unchanged.  ``KeyedColumns`` is this repository's addition: a columnar batch a
source may yield so a whole epoch reaches the CUDA fold without per-item
Python objects.
"""

from __future__ import annotations

import asyncio
import queue
from abc import ABC, abstractmethod
from dataclasses import dataclass
from datetime import datetime, timedelta, timezone
from itertools import islice
from typing import Any, Callable, Generic, Iterable, Iterator, List, Optional, TypeVar

X = TypeVar("X")
S = TypeVar("S")


class AbortExecution(RuntimeError):
    """Raise from ``next_batch`` to stop the run without EOF processing (src/inputs.rs:478-481)."""


class Source(ABC, Generic[X]):  # noqa: B024
    """A location to read items from."""


class StatefulSourcePartition(ABC, Generic[X, S]):
    @abstractmethod
    def next_batch(self) -> Iterable[X]: ...

    def next_awake(self) -> Optional[datetime]:
        return None

    @abstractmethod
    def snapshot(self) -> S: ...

    def close(self) -> None:
        return


class FixedPartitionedSource(Source[X], Generic[X, S]):
    @abstractmethod
    def list_parts(self) -> List[str]: ...

    @abstractmethod
    def build_part(self, step_id: str, for_part: str, resume_state: Optional[S]) -> StatefulSourcePartition[X, S]: ...


class StatelessSourcePartition(ABC, Generic[X]):
    @abstractmethod
    def next_batch(self) -> Iterable[X]: ...

    def next_awake(self) -> Optional[datetime]:
        return None

    def close(self) -> None:
        return


class DynamicSource(Source[X]):
    @abstractmethod
    def build(self, step_id: str, worker_index: int, worker_count: int) -> StatelessSourcePartition[X]: ...


try:  # a defaulted type variable lets `SimplePollingSource[int]` stand for "no resume state" (inputs.py:43-44)
    from typing_extensions import TypeVar as _DefaultedTypeVar

    Sn = _DefaultedTypeVar("Sn", default=None)
except Exception:  # pragma: no cover - typing_extensions too old
    Sn = TypeVar("Sn")


class _SimplePollingPartition(StatefulSourcePartition[X, S]):
    """inputs.py:285-331: one getter call per awake time; the awake times sit on the ``align_to`` grid."""

    def __init__(self, now: datetime, interval: timedelta, align_to: Optional[datetime], getter: Callable[[], Any],
                 snapshot: Callable[[], Any] = lambda: None):
        self._interval, self._getter, self._snapshot = interval, getter, snapshot
        self._next_awake = now
        if align_to is not None:
            past_tick = (now - align_to) % interval
            if past_tick > timedelta(0):  # exactly on a tick: poll right away, not a whole interval later
                self._next_awake = now + (interval - past_tick)

    def next_batch(self):
        try:
            item = self._getter()
            self._next_awake += self._interval
            return [] if item is None else [item]
        except SimplePollingSource.Retry as ex:
            self._next_awake += ex.timeout
            return []

    def next_awake(self):
        return self._next_awake

    def snapshot(self):
        return self._snapshot()


class SimplePollingSource(FixedPartitionedSource[X, Sn]):
    """Call ``next_item`` every ``interval`` on one worker (inputs.py:333-452)."""

    @dataclass
    class Retry(Exception):
        timeout: timedelta

    def __init__(self, interval: timedelta, align_to: Optional[datetime] = None):
        self._interval, self._align_to = interval, align_to

    def list_parts(self):
        return ["singleton"]

    def build_part(self, step_id, for_part, resume_state):
        now = datetime.now(timezone.utc)
        if resume_state is not None:
            self.resume(resume_state)
        return _SimplePollingPartition(now, self._interval, self._align_to, self.next_item, self.snapshot)

    @abstractmethod
    def next_item(self): ...

    def snapshot(self):
        """Position of the next read, handed back to ``resume`` (inputs.py:417-431)."""
        return None

    def resume(self, resume_state) -> None:
        """Called once before ``next_item`` when resuming (inputs.py:433-441)."""
        return None


def batch(ib: Iterable[X], batch_size: int) -> Iterator[List[X]]:
    """Chunk an iterable (inputs.py:455)."""
    it = iter(ib)
    while True:
        chunk = list(islice(it, batch_size))
        if not chunk:
            return
        yield chunk


def batch_getter(getter: Callable[[], X], batch_size: int, yield_on: Optional[X] = None) -> Iterator[List[X]]:
    """Batch from a getter that returns ``yield_on`` when nothing is ready and raises ``StopIteration`` at EOF
    (inputs.py:477): a partial batch is handed out first, then the generator ends."""
    while True:
        chunk, eof = [], False
        while len(chunk) < batch_size:
            try:
                item = getter()
            except StopIteration:
                eof = True
                break
            if item == yield_on:
                break
            chunk.append(item)
        if eof and not chunk:
            return
        yield chunk
        if eof:
            return


def batch_getter_ex(getter: Callable[[], X], batch_size: int, yield_ex=queue.Empty) -> Iterator[List[X]]:
    """Same for a getter that raises ``yield_ex`` (default ``queue.Empty``) when nothing is ready (inputs.py:512)."""
    while True:
        chunk, eof = [], False
        while len(chunk) < batch_size:
            try:
                chunk.append(getter())
            except yield_ex:
                break
            except StopIteration:
                eof = True
                break
        if eof and not chunk:
            return
        yield chunk
        if eof:
            return


def batch_async(aib, timeout: timedelta, batch_size: int, loop=None) -> Iterator[List[X]]:
    """Batch an async iterator with a time limit per batch (inputs.py:546)."""
    loop = loop if loop is not None else asyncio.new_event_loop()
    task = None
    ait = aib.__aiter__()  # an async iterable need not be its own iterator

    async def anext_coro():  # __anext__ may return any awaitable: create_task needs a coroutine
        return await ait.__anext__()

    async def anext_batch():
        nonlocal task
        chunk = []
        deadline = loop.time() + timeout.total_seconds()
        try:
            while len(chunk) < batch_size:
                if task is None:
                    task = loop.create_task(anext_coro())
                remain = deadline - loop.time()
                done, _ = await asyncio.wait({task}, timeout=max(remain, 0))
                if not done:
                    break
                item = task.result()
                task = None
                chunk.append(item)
        except StopAsyncIteration:
            if not chunk:
                raise
        return chunk

    while True:
        try:
            yield loop.run_until_complete(anext_batch())
        except StopAsyncIteration:
            return


@dataclass
class KeyedColumns:
    """A whole batch of ``(key, value)`` items as columns.

    ``keys``: uint64 array (the item key is ``str(key)``), ``ts_us``: int64
    microseconds since the Unix epoch, ``vals``: numeric array or ``None``.
    Stateless operators forward it untouched; a GPU-planned windowing step
    ingests it as one activation.
    """

    keys: Any
    ts_us: Any
    vals: Any = None

    def __len__(self):
        return len(self.keys)
