"""Debug driver: a few C1-shaped activations of configurable size through the streaming path, with stats."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu
A = 1_640_995_200_000_000
B = int(os.environ.get("ROWS", 1 << 20))
NK = int(os.environ.get("NK", 1_000_000))
steps = int(os.environ.get("STEPS", 3))
ctx = gpu.Context(0)
fold = gpu.WindowFold(ctx, "count", 60_000_000, None, A, 0, val_dtype="u64", ts_from_value=True, capacity_hint=NK, max_batch_rows=B, max_emit_rows=1 << 24)
dk, dv = ctx.dev_alloc(B * 8), ctx.dev_alloc(B * 8)
for s in range(steps):
    fold.gen_c1(dk, dv, s * B, B, NK)
    fold.sync()
    t0 = time.perf_counter()
    fold.ingest_device(dk, dv, None, B)
    fold.sync()
    st = fold.stats()
    print(f"step {s}: wall {(time.perf_counter()-t0)*1e3:.2f} ms scatter_ms {st.sum_scatter_ms:.3f} fold_ms {st.sum_fold_ms:.3f} launches {st.kernel_launches} stream {st.combined_folds}", flush=True)
em = fold.advance(); em2 = fold.eof()
print("closed", len(em.closed_key) + len(em2.closed_key), "sum", int(em.closed_acc.sum()) + int(em2.closed_acc.sum()), "expected", steps * B)
