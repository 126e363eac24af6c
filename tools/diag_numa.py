"""Does pinned-buffer placement (NUMA node) matter for H2D on this box?  (diagnostic for bench.py's e2e)"""
import ctypes as C, glob, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bytewax_b200 import gpu

def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out

ctx = gpu.Context(0)
lib = ctx.lib
print("affinity at start:", len(os.sched_getaffinity(0)), "cpus")
print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name'; nvidia-smi topo -m | head -12", shell=True, capture_output=True, text=True).stdout)
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    nodes[p.split("/")[-2]] = cpulist(open(p).read())
gpu_nodes = [open(p).read().strip() for p in glob.glob("/sys/bus/pci/devices/*/numa_node")
             if os.path.exists(os.path.join(os.path.dirname(p), "vendor")) and open(os.path.join(os.path.dirname(p), "vendor")).read().strip() == "0x10de"]
print("numa nodes:", {k: len(v) for k, v in nodes.items()}, "nvidia device numa_node values:", gpu_nodes[:16])
n = 1 << 28
d = ctx.dev_alloc(n)
full = os.sched_getaffinity(0)

def measure(tag):
    h = C.c_void_p()
    lib.bw_host_alloc(ctx.h, n, C.byref(h))
    C.memset(h, 1, n)
    best = 0
    for rep in range(4):
        t = time.perf_counter(); lib.bw_memcpy(ctx.h, C.c_void_p(d), h, n, 0); dt = time.perf_counter() - t
        best = max(best, n / dt / 1e9)
    bd = 0
    for rep in range(3):
        t = time.perf_counter(); lib.bw_memcpy(ctx.h, h, C.c_void_p(d), n, 1); dt = time.perf_counter() - t
        bd = max(bd, n / dt / 1e9)
    print(f"{tag:40s} H2D {best:6.1f} GB/s   D2H {bd:6.1f} GB/s")
    lib.bw_host_free(ctx.h, h)

measure("default affinity")
for name, cpus in nodes.items():
    use = cpus & full
    if not use:
        continue
    os.sched_setaffinity(0, use)
    measure(f"allocated+touched on {name} ({len(use)} cpus)")
os.sched_setaffinity(0, full)
