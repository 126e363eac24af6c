"""Turn an .ncu-rep (read here, no GPU needed) into the per-kernel markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_stream_r2.ncu-rep 16777216 > profiles/r02_kernels_ncu.md
"""
import csv, io, json, os, subprocess, sys

rep, rows_per_launch = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
peak = 6586.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = {
    "t_us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "inst": "smsp__inst_executed.sum",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "grid": "launch__grid_size",
    "block": "launch__block_size", "smem": "launch__shared_mem_per_block_dynamic",
}


def val(r, k):
    i = ix.get(M[k])
    if i is None:
        return None
    v, u = float(r[i]), units[i]
    if u == "Mbyte":
        v *= 1e6
    elif u == "Gbyte":
        v *= 1e9
    elif u == "Kbyte":
        v *= 1e3
    elif u == "ms":
        v *= 1e3
    elif u == "ns":
        v /= 1e3
    return v


agg = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    agg.setdefault(name, []).append(r)
print(f"# ncu summary of `{os.path.basename(rep)}` (`--set full --clock-control none`; HBM peak {peak} GB/s measured)\n")
print(f"Rows per launch of the streaming kernels: {rows_per_launch} (16 algorithmic bytes each, SURVEY 8d).  Durations under ncu are")
print("serialised and cold-cache: the bench's CUDA-event times are the ones quoted elsewhere.\n")
print("| kernel | launches | avg us | DRAM read MB | DRAM write MB | traffic GB/s (frac of peak) | algorithmic GB/s (frac) | warp instr (M) | issue active % | regs | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, rs in agg.items():
    n = len(rs)
    avg = lambda k: sum(val(r, k) or 0 for r in rs) / n  # noqa: E731
    t = avg("t_us")
    traffic = (avg("rd") + avg("wr"))
    tg = traffic / (t * 1e-6) / 1e9 if t else 0
    big = any(s in name for s in ("k_scatter", "k_segfold", "k_fold"))
    ag = 16.0 * rows_per_launch / (t * 1e-6) / 1e9 if (t and big) else None
    print(f"| `{name[:70]}` | {n} | {t:.1f} | {avg('rd')/1e6:.1f} | {avg('wr')/1e6:.1f} | {tg:.0f} ({tg/peak:.3f}) | "
          + (f"{ag:.0f} ({ag/peak:.3f})" if ag else "-") + f" | {avg('inst')/1e6:.2f} | {avg('issue'):.1f} | {avg('regs'):.0f} | {avg('warps'):.1f} |")

# `--traffic out.json`: DRAM bytes per launch of the streaming / fold kernels, the file bench.py reads `roofline.traffic` from
if "--traffic" in sys.argv:
    out_path = sys.argv[sys.argv.index("--traffic") + 1]
    traffic = {}
    if os.path.exists(out_path):
        try:
            old = json.load(open(out_path))
            traffic = {k: v for k, v in old.items() if isinstance(v, dict)}
        except Exception:
            traffic = {}
    for name, rs in agg.items():
        short = name.split("<")[0].strip()
        if short in ("k_scatter", "k_segfold", "k_fold"):
            n = len(rs)
            traffic[short] = {"dram_bytes_per_launch": sum((val(r, "rd") or 0) + (val(r, "wr") or 0) for r in rs) / n,
                              "launches_averaged": n, "source": os.path.basename(rep) + " (ncu --set full, cold cache)"}
    json.dump(traffic, open(out_path, "w"), indent=1)
