#!/usr/bin/env python
"""every utilisation-style metric of an ncu report above a threshold, per kernel:  python tools/ncu_top.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:90], d.get("gpu__time_duration.sum"), "us")
    out = []
    for h, u, v in zip(hdr, units, r):
        if u == "%" or "pct" in h:
            try:
                x = float(v)
            except ValueError:
                continue
            if x >= thr:
                out.append((x, h))
    for x, h in sorted(out, reverse=True):
        print(f"   {x:8.2f}  {h}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            pass
    st = sorted(((float(d[h] or 0), h.split("stalled_")[1].split("_per_")[0]) for h in hdr
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")), reverse=True)[:8]
    print("   stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st))
