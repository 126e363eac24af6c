"""Helpers for testing dataflows: host-side mirror of ``bytewax.testing``."""

from __future__ import annotations

from dataclasses import dataclass
from datetime import datetime, timedelta, timezone
from itertools import islice
from typing import Any, Iterable, Iterator, List, Optional

from bytewax_b200.engine import cluster_main, run_main
from bytewax_b200.inputs import AbortExecution, FixedPartitionedSource, StatefulSourcePartition
from bytewax_b200.outputs import DynamicSink, StatelessSinkPartition

__all__ = ["TestingSink", "TestingSource", "TimeTestingGetter", "ffwd_iter", "poll_next_batch", "run_main", "cluster_main"]


@dataclass
class TimeTestingGetter:
    """A settable clock for unit tests (testing.py:38-60)."""

    now: datetime

    def advance(self, td: timedelta) -> None:
        self.now += td

    def get(self) -> datetime:
        return self.now


def ffwd_iter(it: Iterator[Any], n: int) -> None:
    """Skip ``n`` items of an iterator."""
    next(islice(it, n, n), None)


class _IterSourcePartition(StatefulSourcePartition):
    def __init__(self, ib, batch_size: int, resume_state: Optional[int]):
        self._idx = 0 if resume_state is None else resume_state
        self._batch_size = batch_size
        self._it = iter(ib)
        ffwd_iter(self._it, self._idx)
        self._pending: Optional[BaseException] = None
        self._awake: Optional[datetime] = None

    def next_batch(self) -> List[Any]:
        if self._pending is not None:
            raise self._pending
        self._awake = None
        out: List[Any] = []
        for item in self._it:
            if isinstance(item, TestingSource.EOF):
                self._pending = StopIteration()
                self._idx += 1  # resume after the sentinel
                break
            if isinstance(item, TestingSource.ABORT):
                if not item._triggered:
                    item._triggered = True
                    self._pending = AbortExecution()
                    break
                continue
            if isinstance(item, TestingSource.PAUSE):
                self._awake = datetime.now(timezone.utc) + item.for_duration
                break
            out.append(item)
            if len(out) >= self._batch_size:
                break
        if out or self._pending is not None or self._awake is not None:
            self._idx += len(out)
            return out
        raise StopIteration()

    def next_awake(self):
        return self._awake

    def snapshot(self) -> int:
        return self._idx


class TestingSource(FixedPartitionedSource):
    """Produce the items of an iterable from one worker (testing.py:148-221).

    Sentinels in the iterable: ``EOF()`` ends this execution (a re-run continues
    after it), ``ABORT()`` stops the run abruptly once, ``PAUSE(td)`` holds
    input back for a while.
    """

    __test__ = False

    @dataclass
    class EOF:
        pass

    @dataclass
    class ABORT:
        _triggered: bool = False

    @dataclass
    class PAUSE:
        for_duration: timedelta

    def __init__(self, ib: Iterable[Any], batch_size: int = 1):
        self._ib, self._batch_size = ib, batch_size

    def list_parts(self):
        return ["iterable"]

    def build_part(self, step_id, for_part, resume_state):
        return _IterSourcePartition(self._ib, self._batch_size, resume_state)


class _ListSinkPartition(StatelessSinkPartition):
    def __init__(self, ls: List[Any]):
        self._ls = ls

    def write_batch(self, items: List[Any]) -> None:
        self._ls += items


class TestingSink(DynamicSink):
    """Append every item to a list (testing.py:233-257)."""

    __test__ = False

    def __init__(self, ls: List[Any]):
        self._ls = ls

    def build(self, step_id, worker_index, worker_count):
        return _ListSinkPartition(self._ls)


def poll_next_batch(part, timeout=timedelta(seconds=5)):
    """Call ``next_batch`` until it returns something or the timeout passes (testing.py:260-284)."""
    start = datetime.now(timezone.utc)
    batch = []
    while len(batch) <= 0:
        if datetime.now(timezone.utc) - start > timeout:
            raise TimeoutError()
        batch = part.next_batch()
    return batch


def _main(argv=None) -> None:
    """``python -m bytewax_b200.testing <import_str> -p P -w W``: the reference's local test cluster
    (pysrc/bytewax/testing.py ``__main__``: P processes x W workers over localhost TCP).  Here the P x W workers are
    logical workers of this one process -- state is sharded by key the same way, no sockets."""
    import argparse

    from bytewax_b200.run import _locate_dataflow, _parse_timedelta, _prepare_import

    p = argparse.ArgumentParser(prog="python -m bytewax_b200.testing", description="Test a dataflow on a local cluster")
    p.add_argument("import_str", type=str)
    p.add_argument("-w", "--workers-per-process", type=int, default=1)
    p.add_argument("-p", "--processes", type=int, default=1)
    p.add_argument("-s", "--snapshot-interval", type=_parse_timedelta, default=None)
    p.add_argument("-b", "--backup-interval", type=_parse_timedelta, default=None)
    p.add_argument("-r", "--recovery-directory", default=None)
    args = p.parse_args(argv)
    if args.recovery_directory is not None:
        raise NotImplementedError("recovery is out of scope of this engine (SURVEY.md section 2 row 12)")
    mod_str, _, attr_str = _prepare_import(args.import_str).partition(":")
    flow = _locate_dataflow(mod_str, attr_str)
    cluster_main(flow, [], 0, epoch_interval=args.snapshot_interval,
                 worker_count_per_proc=max(1, args.processes) * max(1, args.workers_per_process))


if __name__ == "__main__":
    _main()
