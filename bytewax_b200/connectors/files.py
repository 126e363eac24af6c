"""Line-oriented file sources and sinks (mirror of ``bytewax.connectors.files``, pysrc/bytewax/connectors/files.py).

Same classes, constructor arguments, partition names (``"<filesystem id>::<path>"`` so that identical files seen by
several workers are read once), resume state (byte offset) and error messages; re-authored.
"""

import csv
import os
from pathlib import Path
from typing import Any, Callable, Dict, Iterator, List, Optional, Union
from zlib import adler32

from bytewax_b200.inputs import FixedPartitionedSource, StatefulSourcePartition, batch
from bytewax_b200.outputs import FixedPartitionedSink, StatefulSinkPartition


def _get_path_dev(path: Path) -> str:
    """Default filesystem id: the device number of the path (files.py:14-15)."""
    return hex(path.stat().st_dev)


def _checked_fs_id(get_fs_id: Callable[[Path], str], where: Path) -> str:
    fs_id = get_fs_id(where)
    if "::" in fs_id:
        raise ValueError(f"result of `get_fs_id` must not contain `::`; got {fs_id!r}")
    return fs_id


def _lines_with_tell(f) -> Iterator[str]:
    """Lines through ``readline`` so that ``f.tell()`` stays usable for snapshots (files.py:18-30)."""
    for line in iter(f.readline, ""):
        yield line


class _FileSourcePartition(StatefulSourcePartition):
    def __init__(self, path: Path, batch_size: int, resume_state: Optional[int]):
        self._f = open(path, "rt")
        if resume_state is not None:
            self._f.seek(resume_state)
        self._batcher = batch((line.rstrip("\n") for line in _lines_with_tell(self._f)), batch_size)

    def next_batch(self) -> List[str]:
        return next(self._batcher)  # StopIteration == EOF

    def snapshot(self) -> int:
        return self._f.tell()

    def close(self) -> None:
        self._f.close()


class DirSource(FixedPartitionedSource):
    """Every file of a directory matching ``glob_pat``, line by line; one partition per file (files.py:58-133)."""

    def __init__(self, dir_path: Path, glob_pat: str = "*", batch_size: int = 1000,
                 get_fs_id: Callable[[Path], str] = _get_path_dev):
        if not dir_path.exists():
            raise ValueError(f"input directory `{dir_path}` does not exist")
        if not dir_path.is_dir():
            raise ValueError(f"input directory `{dir_path}` is not a directory")
        self._dir_path, self._glob_pat, self._batch_size = dir_path, glob_pat, batch_size
        self._fs_id = _checked_fs_id(get_fs_id, dir_path)

    def list_parts(self) -> List[str]:
        if not self._dir_path.exists():
            return []
        return [f"{self._fs_id}::{p.relative_to(self._dir_path)}" for p in self._dir_path.glob(self._glob_pat)]

    def build_part(self, step_id: str, for_part: str, resume_state: Optional[int]):
        _fs_id, rel = for_part.split("::", 1)
        return _FileSourcePartition(self._dir_path / rel, self._batch_size, resume_state)


class FileSource(FixedPartitionedSource):
    """One file, line by line, read by one worker (files.py:136-199)."""

    def __init__(self, path: Union[Path, str], batch_size: int = 1000, get_fs_id: Callable[[Path], str] = _get_path_dev):
        self._path = path if isinstance(path, Path) else Path(path)
        self._batch_size = batch_size
        self._fs_id = _checked_fs_id(get_fs_id, self._path.parent)

    def list_parts(self) -> List[str]:
        return [f"{self._fs_id}::{self._path}"] if self._path.exists() else []

    def build_part(self, step_id: str, for_part: str, resume_state: Optional[int]):
        _fs_id, path = for_part.split("::", 1)
        assert path == str(self._path), "Can't resume reading from different file"
        return _FileSourcePartition(self._path, self._batch_size, resume_state)


class _CSVPartition(StatefulSourcePartition):
    def __init__(self, path: Path, batch_size: int, resume_state: Optional[int], fmtparams: Dict[str, Any]):
        self._f = open(path, "rt", newline="")
        reader = csv.DictReader(_lines_with_tell(self._f), **fmtparams)
        _ = reader.fieldnames  # consume the header before any seek
        if resume_state is not None:
            self._f.seek(resume_state)
        self._batcher = batch(reader, batch_size)

    def next_batch(self) -> List[Dict[str, str]]:
        return next(self._batcher)

    def snapshot(self) -> int:
        return self._f.tell()

    def close(self) -> None:
        self._f.close()


class CSVSource(FixedPartitionedSource):
    """A CSV file as one dict per row, keyed by the header's column names (files.py:231-322)."""

    def __init__(self, path: Path, batch_size: int = 1000, get_fs_id: Callable[[Path], str] = _get_path_dev, **fmtparams):
        self._file_source = FileSource(path, batch_size, get_fs_id)
        self._fmtparams = fmtparams

    def list_parts(self) -> List[str]:
        return self._file_source.list_parts()

    def build_part(self, step_id: str, for_part: str, resume_state: Optional[Any]):
        _fs_id, path = for_part.split("::", 1)
        assert path == str(self._file_source._path), "Can't resume reading from different file"
        return _CSVPartition(self._file_source._path, self._file_source._batch_size, resume_state, self._fmtparams)


class _FileSinkPartition(StatefulSinkPartition):
    def __init__(self, path: Path, resume_state: Optional[int], end: str):
        self._f = open(path, "at")
        self._f.seek(0 if resume_state is None else resume_state)
        self._f.truncate()  # a resumed run rewrites everything after the snapshot offset
        self._end = end

    def write_batch(self, values: List[str]) -> None:
        for v in values:
            self._f.write(v)
            self._f.write(self._end)
        self._f.flush()
        os.fsync(self._f.fileno())

    def snapshot(self) -> int:
        return self._f.tell()

    def close(self) -> None:
        self._f.close()


class DirSink(FixedPartitionedSink):
    """``(key, str_value)`` items spread over ``file_count`` files of a directory by key (files.py:350-416)."""

    def __init__(self, dir_path: Path, file_count: int, file_namer: Callable[[int, int], str] = lambda i, _n: f"part_{i}",
                 assign_file: Callable[[str], int] = lambda k: adler32(k.encode()), end: str = "\n"):
        self._dir_path, self._file_count, self._file_namer, self._assign_file, self._end = dir_path, file_count, file_namer, assign_file, end

    def list_parts(self) -> List[str]:
        return [self._file_namer(i, self._file_count) for i in range(self._file_count)]

    def part_fn(self, item_key: str) -> int:
        return self._assign_file(item_key)

    def build_part(self, step_id: str, for_part: str, resume_state: Optional[int]):
        return _FileSinkPartition(self._dir_path / for_part, resume_state, self._end)


class FileSink(FixedPartitionedSink):
    """``(key, str_value)`` items' values appended to one file, one per line (files.py:419-462)."""

    def __init__(self, path: Union[Path, str], end: str = "\n"):
        self._path = path if isinstance(path, Path) else Path(path)
        self._end = end

    def list_parts(self) -> List[str]:
        return [str(self._path)]

    def part_fn(self, item_key: str) -> int:
        return 0

    def build_part(self, step_id: str, for_part: str, resume_state: Optional[int]):
        return _FileSinkPartition(self._path, resume_state, self._end)
