"""The C oracle (oracle/fold_oracle.c) against the reference-generated goldens and the Python oracle."""

import json
import os
import random

import numpy as np

from oracle import coracle, pyoracle as po


def _run_c(spec, batches, is_float):
    o = coracle.COracle(spec["reduction"], spec["length_us"], spec["offset_us"], spec["align_us"], spec["wait_us"],
                        spec["ordered"], is_float)
    for keys, ts, vals in batches:
        o.on_batch(keys, ts, vals)
    o.on_eof()
    ck, cw, ca, cc, cact = o.closed()
    lk, lw, lv, lts, lact = o.late()
    o.close()
    return (ck, cw, ca, cc, cact), (lk, lw, lv, lts, lact)


def _golden_rows(case, tag):
    rows = []
    for i, act in enumerate(case["acts"]):
        for k, w, t, p in act:
            if t == tag:
                rows.append((k, w, p, i))
    return rows


def test_c_oracle_matches_goldens(golden_dir):
    with open(os.path.join(golden_dir, "window_fold_cases.json")) as f:
        cases = json.load(f)
    for name, case in cases.items():
        s = case["spec"]
        is_float = any(isinstance(v, float) for b in case["batches"] for v in b[2])
        (ck, cw, ca, cc, cact), (lk, lw, lv, lts, lact) = _run_c(s, case["batches"], is_float)
        want_e = _golden_rows(case, "E")
        assert len(ck) == len(want_e), name
        for j, (k, w, p, i) in enumerate(want_e):
            assert (int(ck[j]), int(cw[j]), int(cact[j])) == (k, w, i), (name, j)
            if s["reduction"] == "mean":
                assert ca[j] == p[0] and int(cc[j]) == p[1], (name, j)
            else:
                assert ca[j].item() == p, (name, j, ca[j], p)
        want_l = _golden_rows(case, "L")
        assert len(lk) == len(want_l), name
        for j, (k, w, p, i) in enumerate(want_l):
            assert (int(lk[j]), int(lw[j]), lv[j].item(), int(lact[j])) == (k, w, p, i), (name, j)


def test_c_oracle_vs_python_oracle_random():
    S = 1_000_000
    for seed in range(8):
        rnd = random.Random(seed)
        red = ["count", "sum", "min", "max"][seed % 4]
        spec = po.FoldSpec(red, 10 * S, [None, 5 * S, 3 * S][seed % 3], wait_us=(seed % 3) * S, ordered=bool(seed & 4) and red in ("count", "sum"))
        batches = []
        t0 = spec.align_us
        for b in range(6):
            n = rnd.randint(1, 400)
            keys = [rnd.randrange(13) for _ in range(n)]
            ts = [t0 + b * 8 * S + rnd.randint(-4 * S, 8 * S) for _ in range(n)]
            vals = [rnd.randint(-50, 50) for _ in range(n)]
            batches.append((keys, ts, vals))
        acts = po.run_fold(spec, batches)
        want_e = [(k, w, p, i) for i, a in enumerate(acts) for k, w, t, p in a if t == "E"]
        want_l = [(k, w, p, i) for i, a in enumerate(acts) for k, w, t, p in a if t == "L"]
        d = dict(reduction=red, length_us=spec.length_us, offset_us=spec.offset_us, align_us=spec.align_us, wait_us=spec.wait_us, ordered=spec.ordered)
        (ck, cw, ca, cc, cact), (lk, lw, lv, lts, lact) = _run_c(d, batches, False)
        assert [(int(a), int(b), c.item(), int(e)) for a, b, c, e in zip(ck, cw, ca, cact)] == want_e, seed
        assert [(int(a), int(b), c.item(), int(e)) for a, b, c, e in zip(lk, lw, lv, lact)] == want_l, seed


def test_c1_generator_matches_python():
    keys, ts, vals = coracle.gen_c1(12345, 2000, 1000)
    k2, t2, v2 = po.c1_rows(12345, 2000, 1000)
    assert keys.tolist() == k2 and ts.tolist() == t2 and vals.tolist() == v2


def test_sharded_oracle_equals_single():
    keys, ts, vals = coracle.gen_c1(0, 50_000, 500)
    ts = ts * 1  # us
    single = coracle.COracle("count", 10_000)
    single.on_batch(keys, ts)
    single.on_eof()
    a = single.closed()
    rows = set(zip(a[0].tolist(), a[1].tolist(), a[2].tolist()))
    merged = set()
    for part in range(3):
        o = coracle.COracle("count", 10_000)
        o.on_batch(keys, ts, None, part, 3)
        o.on_eof()
        c = o.closed()
        merged |= set(zip(c[0].tolist(), c[1].tolist(), c[2].tolist()))
        assert all(po.dest_rank(int(k), 3) == part for k in c[0][:50])
    assert rows == merged
