"""Line-oriented file source / sink (subset of pysrc/bytewax/connectors/files.py)."""

import os
from pathlib import Path
from typing import List, Optional, Union

from bytewax_b200.inputs import FixedPartitionedSource, StatefulSourcePartition, batch
from bytewax_b200.outputs import FixedPartitionedSink, StatefulSinkPartition


class _FileSourcePartition(StatefulSourcePartition):
    def __init__(self, path: Path, batch_size: int, resume_state: Optional[int]):
        self._f = open(path, "rt")
        if resume_state is not None:
            self._f.seek(resume_state)
        self._batcher = batch((line.rstrip("\n") for line in iter(self._f.readline, "")), batch_size)

    def next_batch(self):
        return next(self._batcher)  # StopIteration == EOF

    def snapshot(self):
        return self._f.tell()

    def close(self):
        self._f.close()


class FileSource(FixedPartitionedSource):
    """Read a file line by line from one worker."""

    def __init__(self, path: Union[Path, str], batch_size: int = 1000):
        self._path = Path(path)
        self._batch_size = batch_size

    def list_parts(self) -> List[str]:
        return [str(self._path)] if self._path.exists() else []

    def build_part(self, step_id, for_part, resume_state):
        return _FileSourcePartition(self._path, self._batch_size, resume_state)


class _FileSinkPartition(StatefulSinkPartition):
    def __init__(self, path: Path, resume_state: Optional[int], end: str):
        resume_offset = 0 if resume_state is None else resume_state
        self._f = open(path, "at")
        self._f.seek(resume_offset)
        self._f.truncate()
        self._end = end

    def write_batch(self, values):
        for v in values:
            self._f.write(v)
            self._f.write(self._end)
        self._f.flush()
        os.fsync(self._f.fileno())

    def snapshot(self):
        return self._f.tell()

    def close(self):
        self._f.close()


class FileSink(FixedPartitionedSink):
    """Append ``(key, str_value)`` items' values to a file."""

    def __init__(self, path: Union[Path, str], end: str = "\n"):
        self._path = Path(path)
        self._end = end

    def list_parts(self):
        return [str(self._path)]

    def part_fn(self, item_key):
        return 0

    def build_part(self, step_id, for_part, resume_state):
        return _FileSinkPartition(self._path, resume_state, self._end)
