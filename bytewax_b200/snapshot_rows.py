"""Device snapshot -> the reference's recovery rows.

`bw_snapshot_take` (`gpu.WindowFold.snapshot()`) dumps the fold state as columns, one row per live (key, pane).  The reference
snapshots a windowed step per key: `_WindowLogic.snapshot()` returns a `_WindowSnapshot(clock_state, windower_state,
logic_states, queue)` (pysrc/bytewax/operators/windowing.py:1032-1037, 1182-1190) and its recovery store keeps
`(step_id, state_key, epoch, pickle(snapshot))` rows (src/recovery.rs:276-290).  This module builds exactly those objects from
the columns, so that a step that ran on the CUDA path can be resumed by the reference's own Python logic (the builder of
`window()` takes the snapshot as `resume_state`, windowing.py:1287-1303) -- `tests/test_snapshot_rows.py` does that with
the real reference classes.

Panes -> windows: with g = gcd(length, offset), a = offset / g, b = length / g, window w covers panes [w*a, w*a + b); a
window is OPEN for a key iff it covers a live pane and is newer than the key's `closed_upto`; its accumulator is the combine
of its live panes; windows are listed in first-opened order (the earliest arrival index over their panes, then id), which
is the order the reference's dicts have and its close pass emits in.
"""
import math
import pickle
from datetime import datetime, timedelta, timezone
from typing import Any, Dict, List, Tuple

import numpy as np

_EPOCH = datetime(1970, 1, 1, tzinfo=timezone.utc)
_I64_MIN = -(1 << 63)
_UTC_MIN_US = -62_135_596_800_000_000  # datetime.min in UTC


def _dt(us: int) -> datetime:
    return _EPOCH + timedelta(microseconds=int(us))


def _decode(reduction: str, is_float: bool, signed: bool, bits: int):
    """Accumulator bits of the table -> the value the reference's per-window logic holds."""
    bits = int(bits) & 0xFFFFFFFFFFFFFFFF
    if reduction == "count":
        return bits
    if is_float:
        if reduction in ("min", "max"):  # ordered-float encoding (bw_f64_to_ordered)
            bits = (bits & 0x7FFFFFFFFFFFFFFF) if (bits >> 63) else (~bits & 0xFFFFFFFFFFFFFFFF)
        return float(np.array([bits], dtype=np.uint64).view(np.float64)[0])
    if signed and bits >= (1 << 63):
        return bits - (1 << 64)
    return bits


_COMBINE = {"count": lambda x, y: x + y, "sum": lambda x, y: x + y, "min": min, "max": max}


def window_snapshots(snap: dict, *, reduction: str, length_us: int, offset_us: int, align_us: int, wait_us: int, now_us: int = 0,
                     frozen_now_us: int = 0, is_float: bool = False, signed: bool = True, classes=None) -> Dict[str, Any]:
    """`{str(key): _WindowSnapshot}` for every live key of a device snapshot.

    The device keeps one number per key, V = max_j(ts_j - now_j) (now_j: what `bw_fold_set_system_now` said when item j
    arrived); the reference keeps the pair (watermark_base, system_time_of_max_event), of which only the difference
    max_j(ts_j - wait - now_j) matters for every later watermark.  ``now_us`` is the system time to write into the pair
    (the clock reading at the snapshot); base = V - wait + now_us.  A fold that was never told the time (the frozen clock of
    the parity runs, now_j == 0) but whose logic is to resume under a clock frozen at N: pass ``frozen_now_us = N`` (its
    items count as seen at N) and ``now_us = N``.
    ``classes``: the module to take `_WindowSnapshot`, `_EventClockState`, `_SlidingWindowerState`, `WindowMetadata` from
    (default: this package's mirror; the reference's own `bytewax.operators.windowing` gives objects its store can pickle).
    """
    if classes is None:
        from bytewax_b200.operators import windowing as classes
    if reduction not in _COMBINE:
        raise ValueError(f"snapshot rows for reduction {reduction!r} are not supported (count, sum, min, max)")
    g = math.gcd(length_us, offset_us)
    a, b = offset_us // g, length_us // g
    per_key: Dict[int, List[Tuple[int, Any, int]]] = {}
    meta: Dict[int, Tuple[int, int]] = {}
    for k, q, acc, seq, mts, upto in zip(snap["key"].tolist(), snap["pane_id"].tolist(), snap["acc"].tolist(), snap["open_seq"].tolist(),
                                         snap["max_ts_us"].tolist(), snap["closed_upto"].tolist()):
        per_key.setdefault(k, []).append((q, _decode(reduction, is_float, signed, acc), seq))
        meta[k] = (mts, upto)
    comb = _COMBINE[reduction]
    out: Dict[str, Any] = {}
    for k, panes in per_key.items():
        mts, upto = meta[k]
        wins: Dict[int, Tuple[Any, int]] = {}  # window id -> (accumulator, earliest arrival index)
        for q, val, seq in panes:
            w_lo = -((-(q - b + 1)) // a)  # ceil((q - b + 1) / a)
            w_hi = q // a
            for w in range(w_lo, w_hi + 1):
                if upto != _I64_MIN and w <= upto:
                    continue  # already emitted for this incarnation of the key
                cur = wins.get(w)
                wins[w] = (val, seq) if cur is None else (comb(cur[0], val), min(cur[1], seq))
        order = sorted(wins, key=lambda w: (wins[w][1], w))
        opened = {w: classes.WindowMetadata(_dt(align_us + w * offset_us), _dt(align_us + w * offset_us + length_us)) for w in order}
        if mts == _I64_MIN:
            clock = classes._EventClockState(system_time_of_max_event=_dt(now_us), watermark_base=_dt(_UTC_MIN_US))
        else:
            clock = classes._EventClockState(system_time_of_max_event=_dt(now_us),
                                             watermark_base=_dt(max(mts - frozen_now_us + now_us - wait_us, _UTC_MIN_US)))
        out[str(k)] = classes._WindowSnapshot(clock, classes._SlidingWindowerState(opened=opened), {w: wins[w][0] for w in order}, [])
    return out


def recovery_rows(step_id: str, snaps: Dict[str, Any], epoch: int) -> List[Tuple[str, str, int, bytes]]:
    """`(step_id, state_key, snap_epoch, ser_change)` per key: the columns of the reference's `snaps` table
    (src/recovery.rs:276-290, `Upsert` = the pickled snapshot)."""
    return [(step_id, key, int(epoch), pickle.dumps(s)) for key, s in sorted(snaps.items())]
