"""Two loopback ranks (threads) on one GPU fed rank-interleaved, disordered rows: exercises the ROW exchange path
(k_part_hist / k_part_scan / k_part_scatter into the peer's receive region, lateness pass, sort path) so that ncu can
capture those kernels on a single-GPU box.  Diagnostic; not a bench number."""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bytewax_b200 import _native as N, gpu  # noqa: E402

A, S = 1_640_995_200_000_000, 1_000_000
world, n, nb = 2, 1 << 20, 4
ctxs = gpu.Context.loopback_world(world)
rnd = np.random.default_rng(3)
batches = []
for b in range(nb):
    per = []
    for r in range(world):
        ts = A + b * 12 * S + rnd.integers(-6 * S, 12 * S, n)
        per.append((rnd.integers(0, 100_000, n).astype(np.uint64), ts.astype(np.int64), rnd.integers(-100, 100, n)))
    batches.append(per)
done = [None] * world


def rank_main(r):
    fold = gpu.WindowFold(ctxs[r], "sum", 10 * S, 5 * S, A, 2 * S, val_dtype="i64", capacity_hint=1 << 17, max_batch_rows=n,
                          max_emit_rows=1 << 22, max_late_rows=1 << 23, exchange=N.XCHG_P2P)
    for per in batches:
        k, t, v = per[r]
        fold.ingest(k, v, t)
    em, em2 = fold.advance(), fold.eof()
    done[r] = (len(em.closed_key) + len(em2.closed_key), len(em.late_key), fold)


th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
[t.start() for t in th]
[t.join() for t in th]
print("closed / late rows per rank:", [(d[0], d[1]) for d in done])
for d in done:
    d[2].close()
for c in ctxs:
    c.close()
