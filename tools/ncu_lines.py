"""Per-source-line view of an ncu report (needs -lineinfo and --import-source on): instructions executed, stall samples and
active threads per CUDA source line, top N per kernel.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [N] [kernel substring]
"""
import csv
import os
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
kernels = {}
path = func = None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        path = os.path.basename(r[1]); hdr = None; continue
    if len(r) == 2 and r[0] == "Function Name":
        func = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = {}
    for i, h in enumerate(hdr):
        d.setdefault(h, r[i])
    d["file"] = path
    kernels.setdefault(func, []).append(d)
for func, data in kernels.items():
    if want not in func:
        continue
    tot = sum(num(d["Instructions Executed"]) for d in data) or 1
    sam = sum(num(d["# Samples"]) for d in data) or 1
    print(f"== {func}\n   warp instructions {tot}, stall samples {sam}")
    byfile = {}
    for d in data:
        byfile[d["file"]] = byfile.get(d["file"], 0) + num(d["Instructions Executed"])
    print("   by file:", {k: f"{100 * v / tot:.1f}%" for k, v in byfile.items()})
    for d in sorted(data, key=lambda d: -num(d["Instructions Executed"]))[:top_n]:
        stalls = {k[6:]: num(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and num(v) > 0}
        main = ",".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print("%-16s %5s %6.2f%% inst %6.2f%% samp thr %5s  %-40s | %s" % (d["file"][:16], d["Line No"], 100 * num(d["Instructions Executed"]) / tot,
              100 * num(d["# Samples"]) / sam, d.get("Avg. Threads Executed", "")[:5], main[:40], d["Source"].strip()[:90]))
