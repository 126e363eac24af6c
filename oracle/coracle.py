"""ctypes wrapper of oracle/fold_oracle.c.  TEST INFRASTRUCTURE ONLY (see the C file's header)."""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfold_oracle.so")
RED = {"count": 0, "sum": 1, "min": 2, "max": 3, "mean": 4}


def build(force=False):
    src = os.path.join(HERE, "fold_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "_build/libfold_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_on_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        _lib.orc_on_eof.argtypes = [C.c_void_p]
        _lib.orc_n_closed.restype = C.c_size_t
        _lib.orc_n_closed.argtypes = [C.c_void_p]
        _lib.orc_n_late.restype = C.c_size_t
        _lib.orc_n_late.argtypes = [C.c_void_p]
        _lib.orc_closed_col.restype = C.c_void_p
        _lib.orc_closed_col.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_late_col.restype = C.c_void_p
        _lib.orc_late_col.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_clear_rows.argtypes = [C.c_void_p]
        _lib.orc_gen_c1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint64, C.c_int64]
        _lib.orc_on_batch_mt.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    return _lib


class COracle:
    """One worker of the restated reference path over columnar numpy input."""

    def __init__(self, reduction="count", length_us=60_000_000, offset_us=None, align_us=1_640_995_200_000_000,
                 wait_us=0, ordered=False, is_float=False):
        self.l = lib()
        self.is_float = is_float
        self.reduction = reduction
        self.h = C.c_void_p(self.l.orc_create(RED[reduction], int(is_float), length_us, offset_us or length_us,
                                              align_us, wait_us, int(ordered)))

    def on_batch(self, keys, ts, vals=None, part=0, nparts=1):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        ts = np.ascontiguousarray(ts, dtype=np.int64)
        vp = None
        if vals is not None:
            vals = np.ascontiguousarray(vals, dtype=np.float64 if self.is_float else np.int64)
            vp = vals.ctypes.data_as(C.c_void_p)
        self.l.orc_on_batch(self.h, keys.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p), vp, keys.shape[0], part, nparts)

    def on_eof(self):
        self.l.orc_on_eof(self.h)

    def _col(self, fn, n, col, dt):
        if n == 0:
            return np.zeros(0, dtype=dt)
        p = fn(self.h, col)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n,)).view(dt).copy()

    def closed(self):
        """(key, window_id, acc, count, activation) arrays, in emission order."""
        n = self.l.orc_n_closed(self.h)
        acc_dt = np.float64 if (self.is_float and self.reduction != "count") or self.reduction == "mean" else np.int64
        return (self._col(self.l.orc_closed_col, n, 0, np.uint64), self._col(self.l.orc_closed_col, n, 1, np.int64),
                self._col(self.l.orc_closed_col, n, 2, acc_dt), self._col(self.l.orc_closed_col, n, 3, np.uint64),
                self._col(self.l.orc_closed_col, n, 4, np.uint64))

    def late(self):
        n = self.l.orc_n_late(self.h)
        v_dt = np.float64 if self.is_float else np.int64
        return (self._col(self.l.orc_late_col, n, 0, np.uint64), self._col(self.l.orc_late_col, n, 1, np.int64),
                self._col(self.l.orc_late_col, n, 2, v_dt), self._col(self.l.orc_late_col, n, 3, np.int64),
                self._col(self.l.orc_late_col, n, 4, np.uint64))

    def clear(self):
        self.l.orc_clear_rows(self.h)

    def close(self):
        if self.h:
            self.l.orc_destroy(self.h)
            self.h = None


def gen_c1(start, n, n_keys=1_000_000, align_us=1_640_995_200_000_000):
    keys = np.empty(n, np.uint64)
    ts = np.empty(n, np.int64)
    vals = np.empty(n, np.uint64)
    lib().orc_gen_c1(keys.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p),
                     start, n, n_keys, align_us)
    return keys, ts, vals
