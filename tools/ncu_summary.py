#!/usr/bin/env python
"""One-screen summary of an ncu report: per kernel duration, DRAM bytes, issue utilisation, occupancy, stall mix,
and executed-instruction mix per element.   python tools/ncu_summary.py report.ncu-rep [elements_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
elements = float(sys.argv[2]) if len(sys.argv) > 2 else 64 * 4 * 192 * 3136
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:70])
    for k in keys:
        if k in d:
            print(f"   {k:70s} {d[k]}")
    st = sorted(((float(d[h] or 0), h.split("stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)[:7]
    print("   stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
for blk in re.split(r'(?m)^"Kernel Name",', src)[1:]:
    name = blk.splitlines()[0][:60]
    rr = list(csv.reader(io.StringIO("\n".join(blk.splitlines()[1:]))))
    ix = {h: i for i, h in enumerate(rr[0])}
    c = collections.Counter()
    tot = 0
    for r in rr[1:]:
        if not r or not r[0].startswith("0x"):
            continue
        toks = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
        n = float(r[ix["Instructions Executed"]] or 0)
        c[toks[0].split(".")[0]] += n
        tot += n
    per = elements / 32
    print("== instr/element", name, f"total {tot / per:.1f}: " + " ".join(f"{k}={v / per:.2f}" for k, v in c.most_common(24)))
