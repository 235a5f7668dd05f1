#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) into the committed, human-readable evidence under profiles/:

    python tools/make_profile_summary.py gpurun_out/prof_v10.ncu-rep r01 [gpurun_out/launches.csv]

writes profiles/<tag>_ncu_summary.md (per-kernel metrics, stall mix, instruction mix, hottest source lines),
profiles/<tag>_ncu_raw.csv (selected raw metrics) and updates profiles/traffic.json (DRAM bytes per launch, read by
bench.py for roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
out_md = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = [h for h in hdr if any(k in h for k in (
    "Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "launch__registers", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit", "launch__shared_mem_per_block", "sm__warps_active", "smsp__issue_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "sm__throughput", "gpu__dram_throughput", "sm__inst_executed_pipe_xu",
    "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_lsu", "issue_stalled"))]
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_raw.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[hdr.index(k)] for k in keep])
    for r in rows[2:]:
        w.writerow([r[hdr.index(k)] for k in keep])

traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
md = [f"# ncu summary `{tag}` (from `{os.path.basename(rep)}`, `ncu --set full --clock-control none --import-source on`)", "",
      "Captured under the profiler (one launch each, ~40 replay passes): durations here are NOT benchmark numbers; "
      "bench.py times the same kernels with CUDA events.", ""]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"]
    import re as _re
    _m = _re.search(r"(\w+_kernel)", name)
    short = _m.group(1) if _m else name.split("(")[0]
    def g(k):
        return float(d[k]) if d.get(k) not in (None, "") else float("nan")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    rd = g("dram__bytes_read.sum") * scale.get(u["dram__bytes_read.sum"], 1)
    wr = g("dram__bytes_write.sum") * scale.get(u["dram__bytes_write.sum"], 1)
    traffic[short] = rd + wr
    tscale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[u["gpu__time_duration.sum"]]
    dur = g("gpu__time_duration.sum") * tscale
    md += [f"## `{name[:110]}`", "",
           f"* duration under ncu: {dur * 1e6:.1f} us; DRAM read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB "
           f"({(rd + wr) / dur / 1e9:.0f} GB/s under the profiler)",
           f"* registers/thread {d['launch__registers_per_thread']}, grid {d.get('launch__grid_size')}, block {d.get('launch__block_size')}, "
           f"dynamic smem/block {d.get('launch__shared_mem_per_block_dynamic', '?')} {u.get('launch__shared_mem_per_block_dynamic', '')}",
           f"* CTAs/SM limits: registers {d.get('launch__occupancy_limit_registers')}, shared memory {d.get('launch__occupancy_limit_shared_mem')}; "
           f"warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% of 64",
           f"* issue slots busy {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%; executed warp instructions {g('smsp__inst_executed.sum'):.3e}",
           f"* pipes (% of peak while active): fma {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f}, "
           f"alu {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f}, "
           f"xu/MUFU {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f}, lsu {g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f}",
           f"* DRAM throughput {g('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% of ncu's peak; L2 hit rate {g('lts__t_sector_hit_rate.pct'):.1f}%, "
           f"L1 hit rate {g('l1tex__t_sector_hit_rate.pct'):.1f}%; shared-memory bank conflicts "
           f"{g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.3e} of {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.3e} wavefronts"]
    st = sorted(((float(d[h] or 0), h.split("stalled_")[1].split("_per_")[0]) for h in hdr
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")), reverse=True)[:8]
    md += ["* warp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st), ""]
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, short.replace("_kernel", ""), "--top", "14"],
                           capture_output=True, text=True).stdout
    md += ["Hottest source lines (stall samples / executed instructions):", "", "```", lines.rstrip()[:6000], "```", ""]
elems = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--elements=")]      # (b, k, d, l) count of the profiled launch
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep] + elems, capture_output=True, text=True).stdout
md += ["## executed instruction mix per (b, k, d, l) element", "", "```"] + sorted(set(l[:600] for l in summ.splitlines() if l.startswith("== instr/element"))) + ["```", ""]
if launches and os.path.exists(launches):
    md += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised)", "",
           f"see `profiles/{tag}_launches.csv`", ""]
open(out_md, "w").write("\n".join(md))
if "--no-traffic" not in sys.argv:      # traffic.json carries the headline (config-2) workload only
    json.dump(traffic, open(traffic_path, "w"), indent=1)
print(out_md)
