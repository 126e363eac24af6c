"""Debug: multi-window activations through the streaming path vs the C oracle; prints row-level differences."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bytewax_b200 import gpu
from oracle import coracle

S = 1_000_000
A = 1_640_995_200_000_000
n, nb, n_keys = int(os.environ.get("N", 1 << 21)), int(os.environ.get("NB", 3)), int(os.environ.get("NK", 5000))
wait = int(os.environ.get("WAIT", 2))
jit = int(os.environ.get("JIT", 1))
ctx = gpu.Context(0)
fold = gpu.WindowFold(ctx, "count", 10 * S, None, A, wait * S, val_dtype="i64", capacity_hint=8192, max_batch_rows=n, max_emit_rows=1 << 20, max_late_rows=1 << 18)
orc = coracle.COracle("count", 10 * S, None, A, wait * S, False)
rnd = np.random.default_rng(11)
for b in range(nb):
    keys = rnd.integers(0, n_keys, n).astype(np.uint64)
    ts = (A + b * 35 * S + (np.arange(n) * 35 * S) // n + (rnd.integers(-jit * S, jit * S + 1, n) if jit else 0)).astype(np.int64)
    orc.on_batch(keys, ts, np.ones(n, np.int64))
    fold.ingest(keys, None, ts)
orc.on_eof()
em, em_eof = fold.advance(), fold.eof()
ck, cw, ca, _, cact = orc.closed()
gk = np.concatenate([em.closed_key, em_eof.closed_key]); gw = np.concatenate([em.closed_window_id, em_eof.closed_window_id]); ga = np.concatenate([em.closed_acc, em_eof.closed_acc]).astype(np.int64)
ge = np.concatenate([em.closed_epoch, np.full(len(em_eof.closed_key), 99)])
got = collections.Counter(zip(gk.tolist(), gw.tolist(), ga.tolist()))
want = collections.Counter(zip(ck.tolist(), cw.tolist(), ca.tolist()))
print("rows got", len(gk), "want", len(ck), "sum got", ga.sum(), "want", ca.sum())
extra = list((got - want).items())[:12]
missing = list((want - got).items())[:12]
print("extra", extra)
print("missing", missing)
for (k, w, a), _ in extra[:4]:
    print("key", k, "got", [(int(x), int(y), int(e)) for x, y, e, kk in zip(gw, ga, ge, gk) if kk == k], "want", [(int(x), int(y), int(e) + 1) for x, y, e, kk in zip(cw, ca, cact, ck) if kk == k])
st = fold.stats()
print("stats", st.slow_batches, st.fold_launches, st.combined_folds, st.kernel_launches, st.live_keys)
