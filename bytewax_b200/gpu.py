"""Host-side handle on the CUDA windowed-fold path (numpy in, numpy out).

``WindowFold`` is the GPU stand-in for one ``stateful_batch`` step whose logic
is the reference's ``_WindowLogic`` over an ``EventClock`` and a
``SlidingWindower``/``TumblingWindower`` with a numeric fold
(pysrc/bytewax/operators/windowing.py:1046-1190, 1692-1714).  Each
``ingest`` is one activation of src/operators.rs:755-806; ``advance`` returns
what that operator would have given downstream, split like ``WindowOut``
(windowing.py:1193-1222): closed windows (``down`` + ``meta``) and ``late``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _native as N

_NP_VAL = {"u64": np.uint64, "i64": np.int64, "f32": np.float32, "f64": np.float64}


class Context:
    """One rank == one GPU (``bw_ctx``)."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_id: Optional[bytes] = None):
        self.lib = N.load()
        self.rank, self.world, self.device = rank, world, device
        h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        N.check(self.lib.bw_ctx_create(device, rank, world, idbuf, C.byref(h)))
        self.h = h

    @classmethod
    def loopback_world(cls, world: int, device: int = 0) -> "list[Context]":
        """``world`` ranks on ONE device, to be driven by one thread each (``bw_loopback_create``): the multi-rank
        path without a second GPU.  The contexts share the world object; closing the last one frees it."""
        lib = N.load()
        w = C.c_void_p()
        N.check(lib.bw_loopback_create(world, C.byref(w)))
        shared = {"handle": w, "open": world}
        out = []
        for r in range(world):
            c = cls.__new__(cls)
            c.lib, c.rank, c.world, c.device = lib, r, world, device
            h = C.c_void_p()
            N.check(lib.bw_ctx_create_loopback(device, r, w, C.byref(h)))
            c.h = h
            c._loop = shared
            out.append(c)
        return out

    @staticmethod
    def new_nccl_id() -> bytes:
        lib = N.load()
        buf = C.create_string_buffer(128)
        N.check(lib.bw_nccl_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            self.lib.bw_ctx_destroy(self.h)
            self.h = None
            loop = getattr(self, "_loop", None)
            if loop is not None:
                loop["open"] -= 1
                if loop["open"] == 0:
                    self.lib.bw_loopback_destroy(loop["handle"])

    # plain device memory (tests / bench)
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        N.check(self.lib.bw_dev_alloc(self.h, nbytes, C.byref(p)), self.h)
        return p.value

    def dev_free(self, ptr: int):
        N.check(self.lib.bw_dev_free(self.h, C.c_void_p(ptr)), self.h)

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        N.check(self.lib.bw_memcpy(self.h, C.c_void_p(dptr), arr.ctypes.data_as(C.c_void_p), arr.nbytes, 0), self.h)

    def d2h(self, arr: np.ndarray, dptr: int):
        N.check(self.lib.bw_memcpy(self.h, arr.ctypes.data_as(C.c_void_p), C.c_void_p(dptr), arr.nbytes, 1), self.h)

    def flush_l2(self):
        N.check(self.lib.bw_flush_l2(self.h), self.h)


@dataclass
class Emitted:
    """Rows of one ``advance``/``eof`` in the reference's downstream order."""

    closed_key: np.ndarray
    closed_window_id: np.ndarray
    closed_acc: np.ndarray  # typed per reduction / dtype
    closed_count: np.ndarray
    closed_epoch: np.ndarray
    late_key: np.ndarray
    late_window_id: np.ndarray
    late_val: np.ndarray
    late_ts_us: np.ndarray
    late_epoch: np.ndarray

    def down(self, mean: bool = False) -> List[Tuple[int, Tuple[int, object]]]:
        """``(key, (window_id, acc))`` like ``WindowOut.down`` (windowing.py:1196-1200)."""
        if mean:
            return [
                (int(k), (int(w), [float(a), int(c)]))
                for k, w, a, c in zip(self.closed_key, self.closed_window_id, self.closed_acc, self.closed_count)
            ]
        return [(int(k), (int(w), a.item())) for k, w, a in zip(self.closed_key, self.closed_window_id, self.closed_acc)]

    def late(self) -> List[Tuple[int, Tuple[int, object]]]:
        return [(int(k), (int(w), v.item())) for k, w, v in zip(self.late_key, self.late_window_id, self.late_val)]


class WindowFold:
    def __init__(
        self,
        ctx: Context,
        reduction: str = "count",
        length_us: int = 60_000_000,
        offset_us: Optional[int] = None,
        align_to_us: int = 1_640_995_200_000_000,
        wait_us: int = 0,
        val_dtype: str = "u64",
        ts_from_value: bool = False,
        ordered: bool = False,
        emit_order: int = N.ORDER_REFERENCE,
        capacity_hint: int = 1 << 16,
        max_batch_rows: int = 1 << 20,
        max_emit_rows: int = 1 << 20,
        max_late_rows: int = 1 << 16,
        ring_slots: int = 3,
        exchange: int = N.XCHG_P2P,
        final: bool = False,
    ):
        """``final=True``: no event time and no windows -- one accumulator per key, emitted by
        ``eof()`` in key order as window 0 (the ``*_final`` operators); ``ingest`` takes no ``ts``."""
        self.ctx, self.lib = ctx, ctx.lib
        if final:
            ts_from_value, wait_us = False, None
        self.reduction, self.val_dtype, self.ts_from_value = reduction, val_dtype, ts_from_value
        s = N.BwFoldSpec()
        s.struct_size = C.sizeof(N.BwFoldSpec)
        s.reduction = N.RED[reduction]
        s.val_dtype = N.VAL[val_dtype]
        s.ts_source = N.TS_NONE if final else (N.TS_FROM_VALUE if ts_from_value else N.TS_COLUMN)
        s.length_us = length_us
        s.offset_us = offset_us if offset_us is not None else length_us
        s.align_to_us = align_to_us
        s.wait_us = N.BW_WAIT_FOREVER if wait_us is None else wait_us
        s.ordered = int(ordered)
        s.emit_order = emit_order
        s.exchange = exchange
        s.ring_slots = ring_slots
        s.capacity_hint = capacity_hint
        s.max_batch_rows = max_batch_rows
        s.max_emit_rows = max_emit_rows
        s.max_late_rows = max_late_rows
        self.spec = s
        self.has_ts = not ts_from_value and not final
        self.has_vals = True
        h = C.c_void_p()
        N.check(self.lib.bw_fold_create(ctx.h, C.byref(s), C.byref(h)), ctx.h)
        self.h = h
        self._epoch = 0

    # -- ingest ------------------------------------------------------------
    def acquire(self, rows: Optional[int] = None) -> N.BwBatch:
        b = N.BwBatch()
        n = self.spec.max_batch_rows if rows is None else rows
        N.check(self.lib.bw_ingest_acquire(self.h, n, C.byref(b)), self.ctx.h)
        return b

    def slot_arrays(self, b: N.BwBatch):
        """numpy views on a borrowed pinned slot."""
        cap = int(b.capacity)
        keys = np.ctypeslib.as_array(b.keys, shape=(cap,))
        vals = None
        if b.vals:
            vals = np.ctypeslib.as_array(C.cast(b.vals, C.POINTER(C.c_uint8)), shape=(cap * np.dtype(_NP_VAL[self.val_dtype]).itemsize,)).view(_NP_VAL[self.val_dtype])
        ts = np.ctypeslib.as_array(b.ts_us, shape=(cap,)) if b.ts_us else None
        return keys, vals, ts

    def commit(self, b: N.BwBatch, rows: int, epoch: Optional[int] = None):
        if epoch is None:
            self._epoch += 1
            epoch = self._epoch
        N.check(self.lib.bw_ingest_commit(self.h, C.byref(b), rows, epoch), self.ctx.h)

    def ingest(self, keys, vals=None, ts=None, epoch: Optional[int] = None):
        """One activation from host arrays (pinned slot -> async H2D -> kernels)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = keys.shape[0]
        b = self.acquire(n)
        k, v, t = self.slot_arrays(b)
        k[:n] = keys
        if self.has_vals:
            if vals is None:
                if self.reduction != "count":
                    raise ValueError("this fold needs a value column")
                v[:n] = 0
            else:
                v[:n] = np.asarray(vals).astype(_NP_VAL[self.val_dtype], copy=False)
        if self.has_ts:
            if ts is None:
                raise ValueError("this fold needs a ts_us column")
            t[:n] = np.asarray(ts, dtype=np.int64)
        self.commit(b, n, epoch)

    def ingest_device(self, d_keys: int, d_vals: Optional[int], d_ts: Optional[int], rows: int, epoch: Optional[int] = None):
        if epoch is None:
            self._epoch += 1
            epoch = self._epoch
        N.check(
            self.lib.bw_ingest_device(self.h, C.c_void_p(d_keys), C.c_void_p(d_vals or 0), C.c_void_p(d_ts or 0), rows, epoch),
            self.ctx.h,
        )

    def gen_c1(self, d_keys: int, d_vals: int, start: int, rows: int, n_keys: int):
        N.check(self.lib.bw_gen_c1(self.h, C.c_void_p(d_keys), C.c_void_p(d_vals), start, rows, n_keys), self.ctx.h)

    # -- results -----------------------------------------------------------
    def _acc_dtype(self):
        if self.reduction == "count":
            return np.uint64
        if self.reduction == "mean":
            return np.float64
        if self.val_dtype in ("f32", "f64"):
            return np.float64
        return np.int64 if self.val_dtype == "i64" else np.uint64

    def _wrap(self, e: N.BwEmit, copy: bool = True) -> Emitted:
        nc, nl = int(e.n_closed), int(e.n_late)

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            a = np.ctypeslib.as_array(ptr, shape=(n,)).view(dt)
            return a.copy() if copy else a

        late_dt = np.float64 if self.val_dtype in ("f32", "f64") else (np.int64 if self.val_dtype == "i64" else np.uint64)
        return Emitted(
            arr(e.closed_key, nc, np.uint64), arr(e.closed_window_id, nc, np.int64), arr(e.closed_acc, nc, self._acc_dtype()),
            arr(e.closed_count, nc, np.uint64), arr(e.closed_epoch, nc, np.uint64),
            arr(e.late_key, nl, np.uint64), arr(e.late_window_id, nl, np.int64), arr(e.late_val, nl, late_dt),
            arr(e.late_ts_us, nl, np.int64), arr(e.late_epoch, nl, np.uint64),
        )

    def set_system_now(self, now_us: int):
        """System time of the activations ingested from now on (``bw_fold_set_system_now``; never set: frozen clock)."""
        N.check(self.lib.bw_fold_set_system_now(self.h, int(now_us)), self.ctx.h)

    def advance(self, copy: bool = True, system_now_us: int = 0) -> Emitted:
        """Rows emitted since the last call.  ``copy=False`` returns views of the library's pinned
        output buffers (the C ABI's own contract: valid until the next ``advance``/``eof``).
        ``system_now_us`` > 0 also runs the notify phase at that system time (idle keys' windows close)."""
        e = N.BwEmit()
        N.check(self.lib.bw_advance(self.h, 0, int(system_now_us), C.byref(e)), self.ctx.h)
        return self._wrap(e, copy)

    def eof(self, copy: bool = True) -> Emitted:
        e = N.BwEmit()
        N.check(self.lib.bw_eof(self.h, C.byref(e)), self.ctx.h)
        return self._wrap(e, copy)

    _SNAP_COLS = (("key", np.uint64), ("pane_id", np.int64), ("acc", np.uint64), ("count", np.uint64), ("open_seq", np.uint64),
                  ("max_ts_us", np.int64), ("closed_upto", np.int64))

    def snapshot(self) -> dict:
        """State after the last ``advance`` as numpy columns (one row per live (key, pane)) plus three scalars:
        the columnar form of ``_WindowLogic.snapshot`` (windowing.py:1182-1190) for every key at once."""
        sn = N.BwSnapshot()
        N.check(self.lib.bw_snapshot_take(self.h, C.byref(sn)), self.ctx.h)
        n = int(sn.n)
        out = {name: (np.ctypeslib.as_array(getattr(sn, name), shape=(n,)).view(dt).copy() if n else np.zeros(0, dt))
               for name, dt in self._SNAP_COLS}
        out.update(batch_no=int(sn.batch_no), gmax_ts_us=int(sn.gmax_ts_us), last_epoch=int(sn.last_epoch))
        return out

    def restore(self, snap: dict):
        """Load ``snapshot()`` output into this freshly created fold (same window spec; any capacity / world size)."""
        sn = N.BwSnapshot()
        keep = []
        n = len(snap["key"])
        sn.n = n
        for name, dt in self._SNAP_COLS:
            a = np.ascontiguousarray(snap[name], dtype=dt)
            keep.append(a)
            ctype = C.c_int64 if dt is np.int64 else C.c_uint64
            setattr(sn, name, a.ctypes.data_as(C.POINTER(ctype)))
        sn.batch_no, sn.gmax_ts_us, sn.last_epoch = snap["batch_no"], snap["gmax_ts_us"], snap["last_epoch"]
        N.check(self.lib.bw_snapshot_load(self.h, C.byref(sn)), self.ctx.h)
        self._epoch = max(self._epoch, int(snap["last_epoch"]))

    def window_bounds(self, window_id: int) -> Tuple[int, int]:
        o, c = C.c_int64(), C.c_int64()
        self.lib.bw_window_bounds(C.byref(self.spec), window_id, C.byref(o), C.byref(c))
        return o.value, c.value

    def stats(self) -> N.BwStats:
        st = N.BwStats()
        N.check(self.lib.bw_fold_stats(self.h, C.byref(st)), self.ctx.h)
        return st

    def reset_timers(self):
        N.check(self.lib.bw_fold_reset_timers(self.h), self.ctx.h)

    def time_begin(self):
        N.check(self.lib.bw_fold_time_begin(self.h), self.ctx.h)

    def time_end(self) -> float:
        ms = C.c_float()
        N.check(self.lib.bw_fold_time_end(self.h, C.byref(ms)), self.ctx.h)
        return ms.value

    def sync(self):
        N.check(self.lib.bw_fold_sync(self.h), self.ctx.h)

    def close(self):
        if self.h:
            self.lib.bw_fold_destroy(self.h)
            self.h = None


class ZScoreMap:
    """``stateful_map`` with the rolling z-score detector (``bw_smap``; examples/anomaly_detector.py:16-48)."""

    def __init__(self, ctx: Context, window: int = 10, threshold: float = 2.0, val_dtype: str = "f32", capacity_hint: int = 1 << 16,
                 max_batch_rows: int = 1 << 20):
        self.ctx, self.lib, self.val_dtype = ctx, ctx.lib, val_dtype
        s = N.BwSmapSpec()
        s.struct_size = C.sizeof(N.BwSmapSpec)
        s.window, s.val_dtype, s.threshold = window, N.VAL[val_dtype], threshold
        s.capacity_hint, s.max_batch_rows = capacity_hint, max_batch_rows
        h = C.c_void_p()
        N.check(self.lib.bw_smap_create(ctx.h, C.byref(s), C.byref(h)), ctx.h)
        self.h = h

    def apply(self, keys, vals):
        """One activation -> (mu f64[n], sigma f64[n], anomalous bool[n]) aligned with the input rows."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=_NP_VAL[self.val_dtype])
        n = keys.shape[0]
        mu, sigma, flag = np.empty(n, np.float64), np.empty(n, np.float64), np.empty(n, np.uint8)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        N.check(self.lib.bw_smap_apply(self.h, vp(keys), vp(vals), n, vp(mu), vp(sigma), vp(flag)), self.ctx.h)
        return mu, sigma, flag.astype(bool)

    def close(self):
        if self.h:
            self.lib.bw_smap_destroy(self.h)
            self.h = None


class KeyedJoin:
    """Two-sided keyed join (``bw_join``; operators/__init__.py:2157-2190)."""

    INSERT = {"first": 0, "last": 1}
    EMIT = {"complete": 0, "final": 1, "running": 2}

    def __init__(self, ctx: Context, insert_mode: str = "last", emit_mode: str = "complete", capacity_hint: int = 1 << 16,
                 max_batch_rows: int = 1 << 20, max_emit_rows: int = 1 << 20):
        self.ctx, self.lib = ctx, ctx.lib
        s = N.BwJoinSpec()
        s.struct_size = C.sizeof(N.BwJoinSpec)
        s.insert_mode, s.emit_mode = self.INSERT[insert_mode], self.EMIT[emit_mode]
        s.capacity_hint, s.max_batch_rows, s.max_emit_rows = capacity_hint, max_batch_rows, max_emit_rows
        h = C.c_void_p()
        N.check(self.lib.bw_join_create(ctx.h, C.byref(s), C.byref(h)), ctx.h)
        self.h = h
        self._epoch = 0

    def apply(self, keys, sides, vals, epoch: Optional[int] = None):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        sides = np.ascontiguousarray(sides, dtype=np.uint8)
        vals = np.ascontiguousarray(vals, dtype=np.uint64)
        if epoch is None:
            self._epoch += 1
            epoch = self._epoch
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        N.check(self.lib.bw_join_apply(self.h, vp(keys), vp(sides), vp(vals), keys.shape[0], epoch), self.ctx.h)

    def _rows(self, r: N.BwJoinRows):
        n = int(r.n)

        def arr(ptr):
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.uint64)

        key, l, rr, mask, epoch = arr(r.key), arr(r.left), arr(r.right), arr(r.mask), arr(r.epoch)
        rows = [(int(k), int(a) if m & 1 else None, int(b) if m & 2 else None) for k, a, b, m in zip(key, l, rr, mask)]
        return rows, epoch

    def advance(self):
        """Rows ``(key, left_or_None, right_or_None)`` emitted since the last call, in the reference's order."""
        r = N.BwJoinRows()
        N.check(self.lib.bw_join_advance(self.h, C.byref(r)), self.ctx.h)
        return self._rows(r)

    def eof(self):
        r = N.BwJoinRows()
        N.check(self.lib.bw_join_eof(self.h, C.byref(r)), self.ctx.h)
        return self._rows(r)

    def close(self):
        if self.h:
            self.lib.bw_join_destroy(self.h)
            self.h = None
