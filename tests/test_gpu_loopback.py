"""Multi-rank parity on ONE GPU: W ranks as threads of this process over a loopback world (include/bwgpu.h,
bw_loopback_create).  The routing hash, the partition kernels, the exchange into peer receive regions, the
combine-at-source / merge-at-owner pair and the verdict gather all run exactly as they do over NCCL + CUDA IPC
(tests/multi_gpu_worker.py covers that on a multi-GPU box); only the collective's transport differs.

Per destination rank the rows arrive as: source rank 0's rows of the activation, then source rank 1's, ... -- the order
the exchange guarantees -- and that stream is replayed through the C oracle and compared row for row.
Reference: src/timely.rs:455-569 (routed exchange), src/timely.rs:809-815 (the pact's hash route).
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A = 1_640_995_200_000_000
S = 1_000_000


def make_batches(world, case, n=30_000, nb=5, n_keys=3000, seed=123):
    rnd = np.random.default_rng(seed)
    out = []
    for b in range(nb):
        per_rank = []
        for r in range(world):
            if case == "inorder":
                base = A + (b * world + r) * 10 * S
                ts = base + (np.arange(n) * 10 * S) // n
            else:
                ts = A + b * 12 * S + rnd.integers(-6 * S, 12 * S, n)
            keys = rnd.integers(0, n_keys, n).astype(np.uint64) * np.uint64(7919)
            vals = rnd.integers(-100, 100, n)
            per_rank.append((keys, ts.astype(np.int64), vals.astype(np.int64)))
        out.append(per_rank)
    return out


def run_world(world, red, length, offset, wait, batches, val_dtype="i64", **kw):
    from bytewax_b200 import _native as N, gpu

    ctxs = gpu.Context.loopback_world(world)
    results, errors = [None] * world, [None] * world

    def rank_main(r):
        try:
            fold = gpu.WindowFold(ctxs[r], red, length * S, offset * S if offset else None, A, wait * S, val_dtype=val_dtype,
                                  capacity_hint=8192, max_batch_rows=1 << 16, max_emit_rows=1 << 20, max_late_rows=1 << 21,
                                  exchange=N.XCHG_P2P, **kw)
            for per_rank in batches:
                k, t, v = per_rank[r]
                fold.ingest(k, v, t)
            em, em2 = fold.advance(), fold.eof()
            st = fold.stats()
            results[r] = dict(
                ck=np.concatenate([em.closed_key, em2.closed_key]), cw=np.concatenate([em.closed_window_id, em2.closed_window_id]),
                ca=np.concatenate([em.closed_acc, em2.closed_acc]), lk=np.concatenate([em.late_key, em2.late_key]),
                lw=np.concatenate([em.late_window_id, em2.late_window_id]), lv=np.concatenate([em.late_val, em2.late_val]),
                slow=int(st.slow_batches), stream=int(st.combined_folds), fold=fold)
        except BaseException as ex:  # noqa: BLE001 - reported by the main thread
            errors[r] = ex

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    # every rank is done with its peers' memory: only now free anything
    for res in results:
        if res is not None:
            res.pop("fold").close()
    for c in ctxs:
        c.close()
    for ex in errors:
        if ex is not None:
            raise ex
    return results


def oracle_rows(world, d, red, length, offset, wait, batches):
    from oracle import coracle, pyoracle as po

    orc = coracle.COracle(red, length * S, offset * S if offset else None, A, wait * S)
    for per_rank in batches:
        ks, tss, vs = [], [], []
        for r in range(world):
            k, t, v = per_rank[r]
            m = np.fromiter((po.dest_rank(int(x), world) == d for x in k), dtype=bool, count=len(k))
            ks.append(k[m]); tss.append(t[m]); vs.append(v[m])
        orc.on_batch(np.concatenate(ks), np.concatenate(tss), np.concatenate(vs))
    orc.on_eof()
    return orc.closed(), orc.late()


CASES = [
    ("inorder", "count", 10, None, 0),
    ("inorder", "sum", 10, None, 0),
    ("inorder", "min", 10, None, 0),
    ("inorder", "max", 10, 5, 0),
    ("inorder", "mean", 10, None, 0),
    ("disorder", "sum", 10, 5, 2),
    ("disorder", "count", 10, None, 0),
]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case,red,length,offset,wait", CASES)
def test_loopback_world_matches_oracle(world, case, red, length, offset, wait, monkeypatch):
    batches = make_batches(world, case)
    got = run_world(world, red, length, offset, wait, batches)
    total_closed = 0
    for d in range(world):
        (ck, cw, ca, _, _), (lk, lw, lv, _, _) = oracle_rows(world, d, red, length, offset, wait, batches)
        g = got[d]
        assert g["ck"].tolist() == ck.tolist(), (d, len(ck), len(g["ck"]))
        assert g["cw"].tolist() == cw.tolist()
        if red == "mean":
            assert np.array_equal(np.asarray(g["ca"], dtype=np.float64), np.asarray(ca, dtype=np.float64))
        else:
            assert np.asarray(g["ca"]).astype(np.int64).tolist() == np.asarray(ca).astype(np.int64).tolist()
        assert g["lk"].tolist() == lk.tolist() and g["lw"].tolist() == lw.tolist()
        assert np.asarray(g["lv"]).astype(np.int64).tolist() == np.asarray(lv).astype(np.int64).tolist()
        total_closed += len(ck)
        if case == "inorder":
            assert g["slow"] == 0
    assert total_closed > 1000  # not vacuous
    # every rank owns some keys: the route really spread them
    assert all(len(g["ck"]) > 0 for g in got)


def test_loopback_inorder_takes_combiner_path():
    """In-order slices over peer memory: combine at the source, 32-byte partials into the owner's receive region, merge
    there -- one collective per activation (multi-GPU stream path)."""
    import os

    if os.environ.get("BW_STREAM", "1") == "0":
        pytest.skip("streaming path disabled by BW_STREAM=0")
    world = 2
    batches = make_batches(world, "inorder")
    got = run_world(world, "sum", 10, None, 0, batches)
    for g in got:
        assert g["stream"] == len(batches), g["stream"]


@pytest.mark.parametrize("red,offset", [("count", None), ("sum", 5)])
def test_loopback_world_direct_fold(red, offset, monkeypatch):
    """The same in-order slices with the streaming stage off (BW_STREAM=0): row exchange into the peer's receive
    region + the direct table fold, the round-1 path that the not-clean activations still take."""
    monkeypatch.setenv("BW_STREAM", "0")
    world = 2
    batches = make_batches(world, "inorder")
    got = run_world(world, red, 10, offset, 0, batches)
    for d in range(world):
        (ck, cw, ca, _, _), (lk, _, _, _, _) = oracle_rows(world, d, red, 10, offset, 0, batches)
        g = got[d]
        assert g["ck"].tolist() == ck.tolist() and g["cw"].tolist() == cw.tolist()
        assert np.asarray(g["ca"]).astype(np.int64).tolist() == np.asarray(ca).astype(np.int64).tolist()
        assert g["lk"].tolist() == lk.tolist()
        assert g["stream"] == 0 and g["slow"] == 0


def test_loopback_rejects_nccl_exchange():
    from bytewax_b200 import _native as N, gpu

    ctxs = gpu.Context.loopback_world(2)
    try:
        with pytest.raises(N.BwError):
            gpu.WindowFold(ctxs[0], "sum", 10 * S, None, A, 0, val_dtype="i64", capacity_hint=1024, max_batch_rows=1024,
                           exchange=N.XCHG_NCCL)
    finally:
        for c in ctxs:
            c.close()
