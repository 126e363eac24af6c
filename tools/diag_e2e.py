"""Per-step timing of the host-ingest path (diagnostic)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu
A = 1_640_995_200_000_000
B = 1 << 24
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = gpu.Context(0)
fold = gpu.WindowFold(ctx, "count", 60_000_000, None, A, 0, val_dtype="u64", ts_from_value=True, capacity_hint=1_000_000,
                      max_batch_rows=B, max_emit_rows=1 << 25, ring_slots=K + 1)
dk, dv = ctx.dev_alloc(B * 8), ctx.dev_alloc(B * 8)
slots = []
for s in range(K):
    b = fold.acquire(B)
    fold.gen_c1(dk, dv, s * B, B, 1_000_000); fold.sync()
    ctx.lib.bw_memcpy(ctx.h, C.cast(b.keys, C.c_void_p), C.c_void_p(dk), B * 8, 1)
    ctx.lib.bw_memcpy(ctx.h, b.vals, C.c_void_p(dv), B * 8, 1)
    slots.append(b)
t0 = time.perf_counter()
for s in range(K):
    t = time.perf_counter()
    fold.commit(slots[s], B)
    print(f"commit {s}: {1e3*(time.perf_counter()-t):.2f} ms")
fold.sync()
print(f"total {1e3*(time.perf_counter()-t0):.1f} ms for {K} steps -> {K*B/(time.perf_counter()-t0)/1e9:.2f} G ev/s")
t = time.perf_counter(); em = fold.advance(copy=False); print(f"advance {1e3*(time.perf_counter()-t):.1f} ms rows {len(em.closed_key)}")
t = time.perf_counter(); em = fold.eof(copy=False); print(f"eof {1e3*(time.perf_counter()-t):.1f} ms rows {len(em.closed_key)}")
