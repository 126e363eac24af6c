"""Diagnostic: time of the scatter stage alone under BW_SC_DBG variants (results are wrong by construction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu
A = 1_640_995_200_000_000
B = 1 << 24
ctx = gpu.Context(0)
fold = gpu.WindowFold(ctx, "count", 60_000_000, None, A, 0, val_dtype="u64", ts_from_value=True, capacity_hint=1_000_000, max_batch_rows=B, max_emit_rows=1 << 24)
dk, dv = ctx.dev_alloc(B * 8), ctx.dev_alloc(B * 8)
for s in range(6):
    fold.gen_c1(dk, dv, s * B, B, 1_000_000)
    fold.sync()
    fold.ingest_device(dk, dv, None, B)
    fold.sync()
st = fold.stats()
print(os.environ.get("BW_SC_DBG", "0"), "scatter+verdict avg ms", st.sum_scatter_ms / max(1, st.scatter_launches), "fold avg ms", st.sum_fold_ms / max(1, st.timed_folds))
