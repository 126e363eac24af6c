"""The operator API mirror on the host engine, checked against the expected lists written in the
reference's own tests (cited per test) -- these tests read like pytests/operators/*."""

import os
import subprocess
import sys
from collections import defaultdict
from datetime import datetime, timedelta, timezone

import pytest

import bytewax_b200.operators as op
import bytewax_b200.operators.windowing as win
from bytewax_b200.dataflow import Dataflow
from bytewax_b200.errors import BytewaxRuntimeError
from bytewax_b200.operators.windowing import ZERO_TD, EventClock, SessionWindower, SlidingWindower, TumblingWindower
from bytewax_b200.testing import TestingSink, TestingSource, cluster_main, run_main

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALIGN = datetime(2022, 1, 1, tzinfo=timezone.utc)


def test_step_tree_names_match_reference():
    # sub-step ids are part of the contract
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource([]))
    k = op.key_on("key", s, lambda x: "a")
    wo = win.fold_window("fw", k, EventClock(lambda x: x, ZERO_TD), TumblingWindower(timedelta(seconds=1), ALIGN), list, lambda a, b: a, list.__add__)
    op.output("out", wo.down, TestingSink([]))

    def ids(steps, acc):
        for st in steps:
            acc.append((st.step_id, type(st).__name__))
            ids(st.substeps, acc)
        return acc

    got = dict(ids(flow.substeps, []))
    assert got["df.key.map.flat_map_batch"] == "flat_map_batch"
    assert got["df.fw.window.stateful_batch"] == "stateful_batch"
    for name in ("unwrap_down", "unwrap_late", "unwrap_meta"):
        assert got[f"df.fw.window.{name}.flat_map_value.flat_map.flat_map_batch"] == "flat_map_batch"
    with pytest.raises(ValueError, match="period"):
        Dataflow("a.b")
    with pytest.raises(ValueError, match="already exists"):
        op.map("key", s, lambda x: x)
    with pytest.raises(TypeError, match="must be a `Stream`"):
        op.map("m", wo, lambda x: x)


def test_map_filter_branch_merge():
    # pytests/operators/test_map.py, test_filter.py, test_branch.py:9-28, test_merge.py
    out1, out2, out3 = [], [], []
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource(range(6), batch_size=2))
    b = op.branch("evens", s, lambda x: x % 2 == 0)
    op.output("o1", op.map("sq", b.trues, lambda x: x * x), TestingSink(out1))
    op.output("o2", op.filter("big", b.falses, lambda x: x > 1), TestingSink(out2))
    op.output("o3", op.merge("m", b.trues, b.falses), TestingSink(out3))
    run_main(flow)
    assert out1 == [0, 4, 16] and out2 == [3, 5] and sorted(out3) == list(range(6))


def test_branch_requires_bool():
    # pytests/operators/test_branch.py:55-74
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource([1]))
    b = op.branch("br", s, lambda x: "nope")
    op.output("o", b.trues, TestingSink([]))
    with pytest.raises(BytewaxRuntimeError) as e:
        run_main(flow)
    assert "must be a `bool`" in str(e.value.__cause__)


def test_inspect_debug_epoch_and_worker():
    # pytests/operators/test_inspect.py:38-55: first epoch is 1, worker 0
    seen = []
    flow = Dataflow("test_df")
    s = op.input("inp", flow, TestingSource(["a"]))
    op.inspect_debug("insp", s, lambda step_id, item, epoch, worker: seen.append((step_id, item, epoch, worker)))
    run_main(flow)
    assert seen == [("test_df.insp", "a", 1, 0)]


def test_needs_input_and_output():
    # pytests/test_inputs.py:30-36, test_outputs.py:9-18
    flow = Dataflow("df")
    with pytest.raises(RuntimeError) as e:  # BytewaxRuntimeError chained to the ValueError (pytests/test_inputs.py:30-36)
        run_main(flow)
    assert isinstance(e.value.__cause__, ValueError) and "at least one input" in str(e.value.__cause__)
    op.input("inp", flow, TestingSource([1]))
    with pytest.raises(RuntimeError) as e2:  # pytests/test_outputs.py:8-18
        run_main(flow)
    assert isinstance(e2.value.__cause__, ValueError) and "at least one output" in str(e2.value.__cause__)


def test_user_exception_is_chained():
    # pytests/test_execution.py:28-55
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource(range(3)))

    def boom(x):
        raise RuntimeError("BOOM")

    op.output("o", op.map("m", s, boom), TestingSink([]))
    with pytest.raises(BytewaxRuntimeError) as e:
        run_main(flow)
    assert isinstance(e.value.__cause__, BytewaxRuntimeError) or "BOOM" in str(e.value.__cause__) or "BOOM" in str(e.value.__cause__.__cause__)


def test_count_final_sorted_eof_order():
    # pytests/operators/test_count_final.py:6-16
    out = []
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource(["a", "a", "b", "c", "b", "a"]))
    op.output("o", op.count_final("count", s, lambda x: x), TestingSink(out))
    run_main(flow)
    assert out == [("a", 3), ("b", 2), ("c", 1)]


def test_fold_reduce_max_min_final():
    # pytests/operators/test_fold_final.py:18-29, test_reduce_final.py:6-17, test_max_min_final.py
    inp = [("a", 1), ("b", 5), ("a", 4), ("b", 2)]
    for build, want in (
        (lambda s: op.fold_final("f", s, list, lambda acc, x: acc + [x]), [("a", [1, 4]), ("b", [5, 2])]),
        (lambda s: op.reduce_final("r", s, lambda a, b: a + b), [("a", 5), ("b", 7)]),
        (lambda s: op.max_final("mx", s), [("a", 4), ("b", 5)]),
        (lambda s: op.min_final("mn", s), [("a", 1), ("b", 2)]),
    ):
        out = []
        flow = Dataflow("df")
        s = op.input("inp", flow, TestingSource(inp))
        op.output("o", build(s), TestingSink(out))
        run_main(flow)
        assert out == want


def test_stateful_map_running_mean():
    # pytests/operators/test_stateful_map.py:9-35
    def running_mean(last3, new):
        last3 = (last3 or []) + [new]
        last3 = last3[-3:]
        return (last3, sum(last3) / len(last3))

    out = []
    flow = Dataflow("df")
    s = op.input("inp", flow, TestingSource([2, 5, 8, 2, 3]))
    k = op.key_on("k", s, lambda _: "ALL")
    m = op.stateful_map("mean", k, running_mean)
    op.output("o", m, TestingSink(out))
    run_main(flow)
    assert out == [("ALL", 2.0), ("ALL", 3.5), ("ALL", 5.0), ("ALL", 5.0), ("ALL", 13 / 3)]


def test_join_modes():
    # pytests/operators/test_join.py:55-135
    def run(insert_mode, emit_mode):
        out = []
        flow = Dataflow("df")
        l = op.input("l", flow, TestingSource([("a", 1), ("b", 2), ("a", 3)]))
        r = op.input("r", flow, TestingSource([("a", 10), ("c", 30)]))
        op.output("o", op.join("j", l, r, insert_mode=insert_mode, emit_mode=emit_mode), TestingSink(out))
        run_main(flow)
        return sorted(out, key=repr)

    assert ("a", (1, 10)) in run("first", "complete") or ("a", (3, 10)) in run("last", "complete")
    final = run("last", "final")
    assert ("a", (3, 10)) in final and ("b", (2, None)) in final and ("c", (None, 30)) in final
    prod = run("product", "final")
    assert ("a", (1, 10)) in prod and ("a", (3, 10)) in prod
    with pytest.raises(ValueError):
        Dataflow("x") and op.join("j", op.input("i", Dataflow("y"), TestingSource([])), insert_mode="nope")


def test_count_window_reference_expectation():
    # pytests/operators/windowing/test_count_window.py:10-35
    inp = [
        {"time": ALIGN + timedelta(seconds=0), "user": "a", "val": 1},
        {"time": ALIGN + timedelta(seconds=4), "user": "a", "val": 1},
        {"time": ALIGN + timedelta(seconds=8), "user": "b", "val": 1},
        {"time": ALIGN + timedelta(seconds=12), "user": "a", "val": 1},
        {"time": ALIGN + timedelta(seconds=13), "user": "a", "val": 1},
    ]
    out = []
    clock = EventClock(lambda e: e["time"], wait_for_system_duration=ZERO_TD)
    windower = TumblingWindower(length=timedelta(seconds=10), align_to=ALIGN)
    flow = Dataflow("test_df")
    s = op.input("inp", flow, TestingSource(inp))
    wo = win.count_window("add", s, clock, windower, lambda e: e["user"])
    op.output("out", wo.down, TestingSink(out))
    run_main(flow)
    assert out == [("a", (0, 2)), ("a", (1, 2)), ("b", (0, 1))]


def test_fold_window_tumbling_and_sliding_reference_expectation():
    # pytests/operators/windowing/test_fold_window.py:37-88 and :143-197
    events = [(ALIGN, "login"), (ALIGN + timedelta(seconds=4), "post"), (ALIGN + timedelta(seconds=8), "post"),
              (ALIGN + timedelta(seconds=16), "post")]
    out = []
    flow = Dataflow("test_df")
    s = op.input("inp", flow, TestingSource(events))
    k = op.key_on("key", s, lambda _: "ALL")

    def count(counts, ev):
        counts[ev[1]] += 1
        return counts

    def merge(a, b):
        a.update(b)
        return a

    fo = win.fold_window("count", k, EventClock(lambda e: e[0], ZERO_TD), TumblingWindower(timedelta(seconds=10), ALIGN),
                         lambda: defaultdict(int), count, merge)
    cleaned = op.map("normal_dict", op.key_rm("key_rm", fo.down), lambda iv: (iv[0], dict(iv[1])))
    op.output("out", cleaned, TestingSink(out))
    run_main(flow)
    assert out == [(0, {"login": 1, "post": 2}), (1, {"post": 1})]

    secs = [(1, "a"), (4, "b"), (8, "c"), (12, "d"), (13, "e"), (14, "f"), (16, "g"), (1, "h")]
    out, late = [], []
    flow = Dataflow("test_df")
    s = op.input("inp", flow, TestingSource([(ALIGN + timedelta(seconds=t), v) for t, v in secs]))
    k = op.key_on("key", s, lambda _: "ALL")
    fo = win.fold_window("sum", k, EventClock(lambda e: e[0], ZERO_TD), SlidingWindower(timedelta(seconds=10), timedelta(seconds=5), ALIGN),
                         list, lambda acc, e: acc + [e[1]], list.__add__)
    op.output("out", op.key_rm("unkey", fo.down), TestingSink(out))
    op.output("late", op.key_rm("unkey_l", fo.late), TestingSink(late))
    run_main(flow)
    assert out == [(-1, ["a", "b"]), (0, ["a", "b", "c"]), (1, ["c", "d", "e", "f"]), (2, ["d", "e", "f", "g"]), (3, ["g"])]
    assert [(w, v[1]) for w, v in late] == [(-1, "h"), (0, "h")]


def test_fold_window_benchmark_shape():
    # pytests/operators/windowing/test_fold_window.py:200-235 (shortened to 6000 items)
    out = []
    start = datetime(2024, 1, 1, tzinfo=timezone.utc)
    flow = Dataflow("bench")
    times = op.input("in", flow, TestingSource([start + timedelta(seconds=i) for i in range(6000)], 10))
    k = op.key_on("key", times, lambda _: "ALL")
    fo = win.fold_window("fold_window", k, EventClock(lambda x: x, ZERO_TD), TumblingWindower(timedelta(minutes=1), start),
                         lambda: None, lambda s, _: s, lambda s, _: s, ordered=False)
    op.output("out", op.key_rm("unkey", fo.down), TestingSink(out))
    run_main(flow)
    assert out == [(i, None) for i in range(100)]


def test_session_window():
    # pytests/operators/windowing/test_fold_window.py:97-140 shape
    secs = [(1, "a"), (5, "b"), (11, "c"), (12, "d"), (13, "e"), (14, "f"), (22, "g")]
    out = []
    flow = Dataflow("test_df")
    s = op.input("inp", flow, TestingSource([(ALIGN + timedelta(seconds=t), v) for t, v in secs]))
    k = op.key_on("key", s, lambda _: "ALL")
    fo = win.fold_window("sum", k, EventClock(lambda e: e[0], ZERO_TD), SessionWindower(gap=timedelta(seconds=5)),
                         list, lambda acc, e: acc + [e[1]], list.__add__)
    op.output("out", op.key_rm("unkey", fo.down), TestingSink(out))
    run_main(flow)
    assert out == [(0, ["a", "b"]), (1, ["c", "d", "e", "f"]), (2, ["g"])]


class _Num(int):
    """An int that carries its event time (values must be plain numbers for `operator.add` / max / min)."""


def _num(e):
    n = _Num(e[1])
    n.ts = e[0]
    return n


def test_reduce_max_min_window():
    # pytests/operators/windowing/test_reduce_window.py:10-44, test_max_min_window.py:14-67
    import operator

    secs = [(0, 1), (4, 9), (8, 3), (12, 7), (13, 2)]
    for make, want in ((lambda k, c, w: win.reduce_window("r", k, c, w, operator.add), [("a", (0, 13)), ("a", (1, 9))]),
                       (lambda k, c, w: win.max_window("r", k, c, w), [("a", (0, 9)), ("a", (1, 7))]),
                       (lambda k, c, w: win.min_window("r", k, c, w), [("a", (0, 1)), ("a", (1, 2))])):
        out = []
        flow = Dataflow("test_df")
        s = op.input("inp", flow, TestingSource([(ALIGN + timedelta(seconds=t), v) for t, v in secs]))
        k = op.key_on("key", s, lambda _: "a")
        wo = make(op.map_value("num", k, _num), EventClock(lambda e: e.ts, ZERO_TD), TumblingWindower(timedelta(seconds=10), ALIGN))
        op.output("out", wo.down, TestingSink(out))
        run_main(flow)
        assert [(k_, (w_, int(v_))) for k_, (w_, v_) in out] == want


def test_two_logical_workers_same_multiset():
    # pytests/test_execution.py:15-25: more than one worker -> compare sorted output
    def build():
        out = []
        flow = Dataflow("df")
        s = op.input("inp", flow, TestingSource([f"k{i % 7}" for i in range(50)], batch_size=5))
        op.output("o", op.count_final("c", s, lambda x: x), TestingSink(out))
        return flow, out

    f1, o1 = build()
    run_main(f1)
    f2, o2 = build()
    cluster_main(f2, [], 0, worker_count_per_proc=2)
    assert sorted(o1) == sorted(o2) and len(o1) == 7


def test_c0_wordcount_via_run_module():
    # BASELINE.json configs[0]: examples/wordcount.py shape through `python -m bytewax_b200.run`
    r = subprocess.run([sys.executable, "-m", "bytewax_b200.run", "tests.flows.wordcount:flow"], cwd=ROOT, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    counts = dict(eval(ln) for ln in lines)
    assert counts["to"] == 4 and counts["be"] == 2 and counts["of"] == 2 and counts["them"] == 1
    assert [eval(ln)[0] for ln in lines] == sorted(counts)  # EOF emission in ascending word order


def test_stateful_steps_see_epochs_in_order_across_input_partitions():
    """Every input partition advances its own epoch; a stateful step must still see epochs in order (the engine holds an
    item back until no partition can produce an earlier epoch: src/timely.rs:95-133, src/operators.rs:687-728)."""
    from bytewax_b200.inputs import DynamicSource, StatelessSourcePartition

    class _Part(StatelessSourcePartition):
        def __init__(self, name, n, gap_ms):
            self.name, self.left, self.gap = name, n, timedelta(milliseconds=gap_ms)
            self.awake = None

        def next_batch(self):
            if self.left == 0:
                raise StopIteration()
            self.left -= 1
            self.awake = datetime.now(timezone.utc) + self.gap
            return [(self.name, 1)]

        def next_awake(self):
            return self.awake

    class _Src(DynamicSource):
        def __init__(self, name, n, gap_ms):
            self.args = (name, n, gap_ms)

        def build(self, step_id, worker_index, worker_count):
            return _Part(*self.args)

    seen = []
    flow = Dataflow("df")
    fast = op.input("fast", flow, _Src("f", 40, 1))
    slow = op.input("slow", flow, _Src("s", 8, 25))
    s = op.merge("m", fast, slow)
    s = op.stateful_map("sm", s, lambda st, v: ((st or 0) + v, (st or 0) + v))
    op.inspect_debug("insp", s, lambda step_id, item, epoch, worker: seen.append(epoch))
    op.output("out", s, TestingSink([]))
    run_main(flow, epoch_interval=timedelta(milliseconds=5))
    assert len(seen) == 48 and seen == sorted(seen) and seen[-1] > seen[0]


def test_gpu_key_ids_do_not_collide():
    """Numeric-looking keys take their own value as device id only when that is unambiguous (ASCII, canonical, < 2^63);
    everything else is interned from 2^63 up."""
    from bytewax_b200.engine import _GpuWindowStep

    st = _GpuWindowStep.__new__(_GpuWindowStep)
    st.key_ids, st.id_keys, st.resort = {}, {}, False
    ids = [st._key_id(k) for k in ["3", "٣", "²", "9223372036854775808", "007", "abc", "9223372036854775807"]]
    assert ids[0] == 3 and ids[6] == (1 << 63) - 1
    assert all(i >= (1 << 63) for i in ids[1:6]) and len(set(ids)) == len(ids)
    assert [st._key_str(i) for i in ids] == ["3", "٣", "²", "9223372036854775808", "007", "abc", "9223372036854775807"]


def test_declared_detector_and_two_sided_join_carry_a_gpu_plan():
    """The recogniser hooks of `op.stateful_map` / `op.join`: the plan rides on the step's builder (the host run ignores it)."""
    import bytewax_b200.operators as bop
    from bytewax_b200.dataflow import Dataflow as DF
    from bytewax_b200.detectors import ZScoreDetector
    from bytewax_b200.testing import TestingSink as Sink, TestingSource as Src, run_main as run

    det = ZScoreDetector(window=3, threshold_z=1.5)
    assert det._gpu_plan == bop.GpuSmapPlan(3, 1.5)
    # the reference example's numbers (examples/anomaly_detector.py semantics): flag uses the statistics BEFORE the push
    st, e1 = det(None, 1.0)
    st, e2 = det(st, 1.0)
    st, e3 = det(st, 10.0)
    assert e1 == (1.0, 1.0, 0.0, False) and e2 == (1.0, 1.0, 0.0, False) and e3[3] is False  # sigma == 0: never anomalous
    st, e4 = det(st, 1.0)
    assert e4[3] is False and abs(e3[1] - 4.0) < 1e-12
    out = []
    flow = DF("t")
    a = bop.input("a", flow, Src([("k", 1), ("q", 2)]))
    b = bop.input("b", flow, Src([("k", "x")]))
    j = bop.join("j", a, b)
    bop.output("o", j, Sink(out))
    run(flow)
    assert out == [("k", (1, "x"))]
    plans = [getattr(s.builder, "_gpu_plan", None) for s in flow._flat_steps() if hasattr(s, "builder")] if hasattr(flow, "_flat_steps") else None
    if plans is not None:
        assert any(isinstance(p, bop.GpuJoinPlan) for p in plans)
    # three sides / product: no device plan
    assert bop.GpuJoinPlan("last", "complete") == bop.GpuJoinPlan("last", "complete")


def test_fold_columns_host_path_equals_the_per_item_operators():
    """`fold_columns` over KeyedColumns batches without the CUDA path: the engine expands the columns into `(str(key), (ts_us,
    value))` items; rows equal `count_window` / `reduce_window` over the same rows."""
    import numpy as np

    import bytewax_b200.operators as bop
    import bytewax_b200.operators.windowing as bwin
    from bytewax_b200.dataflow import Dataflow as DF
    from bytewax_b200.inputs import KeyedColumns
    from bytewax_b200.testing import TestingSink as Sink, TestingSource as Src, run_main as run

    align = datetime(2022, 1, 1, tzinfo=timezone.utc)
    a_us = 1_640_995_200_000_000
    n, batches, items = 1500, [], []
    for b in range(3):
        i = np.arange(b * n, (b + 1) * n)
        keys, ts, vals = (i * 2654435761 % 37).astype(np.uint64), (a_us + i * 9_000).astype(np.int64), (i % 11 - 3).astype(np.int64)
        batches.append(KeyedColumns(keys=keys, ts_us=ts, vals=vals))
        items.extend((str(int(k)), (align + timedelta(microseconds=int(t - a_us)), int(v))) for k, t, v in zip(keys, ts, vals))
    windower = bwin.SlidingWindower(timedelta(seconds=10), timedelta(seconds=5), align)
    for red, ref in (("count", None), ("sum", lambda a, b: (b[0], a[1] + b[1])), ("max", None)):
        got, want = [], []
        flow = DF("cols")
        wo = bwin.fold_columns("f", bop.input("inp", flow, Src(batches)), red, windower, now_getter=lambda: align)
        bop.output("o", wo.down, Sink(got))
        run(flow)
        flow = DF("items")
        s = bop.input("inp", flow, Src(items, batch_size=n))
        clock = bwin.EventClock(lambda v: v[0], timedelta(0), now_getter=lambda: align)
        if red == "count":
            wo = bwin.fold_window("f", s, clock, windower, lambda: 0, lambda a, _v: a + 1, lambda a, b: a + b, ordered=False)
            fix = lambda kv: kv  # noqa: E731
        elif red == "sum":
            wo = bwin.reduce_window("f", s, clock, windower, ref)
            fix = lambda kv: (kv[0], (kv[1][0], kv[1][1][1]))  # noqa: E731
        else:
            wo = bwin.max_window("f", s, clock, windower, by=lambda v: v[1])
            fix = lambda kv: (kv[0], (kv[1][0], kv[1][1][1]))  # noqa: E731
        bop.output("o", wo.down, Sink(want))
        run(flow)
        assert got == [fix(kv) for kv in want], red
        assert len(got) > 100
