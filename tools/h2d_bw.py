"""Raw pinned-host -> device copy bandwidth on this box (context for bench.py's e2e number)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu
ctx = gpu.Context(0)
lib = ctx.lib
n = 1 << 30
h = C.c_void_p(); lib.bw_host_alloc(ctx.h, n, C.byref(h))
C.memset(h, 1, n)
d = ctx.dev_alloc(n)
for rep in range(3):
    t = time.perf_counter(); lib.bw_memcpy(ctx.h, C.c_void_p(d), h, n, 0); dt = time.perf_counter() - t
    print(f"H2D 1 GiB pinned: {n/dt/1e9:.1f} GB/s")
for rep in range(2):
    t = time.perf_counter(); lib.bw_memcpy(ctx.h, h, C.c_void_p(d), n, 1); dt = time.perf_counter() - t
    print(f"D2H 1 GiB pinned: {n/dt/1e9:.1f} GB/s")
