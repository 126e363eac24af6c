"""Per-step fold kernel times for config C1 (diagnostic; not a bench number)."""
import sys, os
os.environ.setdefault("BW_TIMER_STRIDE", "1")  # (per-step numbers: time every activation)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu, _native as N
A = 1_640_995_200_000_000
B = 1 << 24
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = gpu.Context(0)
fold = gpu.WindowFold(ctx, "count", 60_000_000, None, A, 0, val_dtype="u64", ts_from_value=True, capacity_hint=1_000_000,
                      max_batch_rows=B, max_emit_rows=1 << 24)
dk, dv = ctx.dev_alloc(B * 8), ctx.dev_alloc(B * 8)
for s in range(steps):
    fold.gen_c1(dk, dv, s * B, B, 1_000_000)
    fold.sync()
    fold.time_begin()
    fold.ingest_device(dk, dv, None, B)
    ms = fold.time_end()
    st = fold.stats()
    fold_ms = st.sum_fold_ms
    fold.reset_timers()
    print(f"step {s:2d}: step {ms:.3f} ms  fold {fold_ms:.3f} ms  -> {B/fold_ms/1e6:.1f} G ev/s")
em = fold.advance()
print("closed rows", len(em.closed_key))
