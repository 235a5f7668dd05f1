#!/usr/bin/env python
"""Per-source-line summary of an ncu report for one kernel (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep ss2d_fwd [--so xfmamba_b200/libxfscan.so] [--top 40]

ncu's CSV source page is per SASS instruction and carries no line numbers; this joins it (by instruction offset) with
`nvdisasm --print-line-info` of the cubin embedded in the .so (built with -lineinfo) and aggregates executed
instructions and stall samples per CUDA source line.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_line_map(so, kernel_regex):
    """offset -> (file, line, inlined-at chain text) for the first function whose mangled name matches"""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    out = {}
    for cubin in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        cur, fn, loc = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                fn = m.group(1)
                cur = {} if re.search(kernel_regex, fn) and fn not in out else None
                if cur is not None:
                    out[fn] = cur
                continue
            if cur is None:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if m and loc:
                cur[int(m.group(1), 16)] = (loc, m.group(2))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="regex on the kernel name")
    ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xfmamba_b200", "libxfscan.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest SASS instructions")
    ap.add_argument("--fn", default=None, help="regex on the MANGLED cubin function name (default: derived from `kernel`; needed when several instantiations have similar sizes)")
    ap.add_argument("--op", default=None, help="only count SASS instructions whose opcode matches this regex (e.g. 'IMAD.MOV|^MOV')")
    a = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}"],
                         capture_output=True, text=True).stdout
    blocks = re.split(r'(?m)^"Kernel Name",', raw)
    if len(blocks) < 2:
        sys.exit("kernel not found in report")
    first = blocks[1]
    kname = first.splitlines()[0]
    rows = list(csv.reader(io.StringIO("\n".join(first.splitlines()[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) > ix["# Samples"] and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    mangled_hint = a.fn or re.sub(r"\W+", ".*", a.kernel)
    maps = sass_line_map(a.so, mangled_hint)
    # choose the function whose SASS matches the report: same instruction count, then most identical opcodes
    def score(fn):
        mp_ = maps[fn]
        if len(mp_) != len(data):
            return -abs(len(mp_) - len(data))
        same = 0
        for r in data:
            ent = mp_.get(int(r[0], 16) - base)
            if ent and ent[1].split()[:2] == r[ix["Source"]].split()[:2]:
                same += 1
        return same
    best = max(maps, key=score)
    mp = maps[best]
    print(f"# kernel: {kname.strip(',')[:120]}\n# sass function: {best} ({len(mp)} instr in cubin, {len(data)} in report)")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per_line = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    tot_i = tot_s = 0.0
    sass_rows = []
    for r in data:
        off = int(r[0], 16) - base
        loc, _ = mp.get(off, ((("?", 0, "")), ""))
        inst = float(r[ix["Instructions Executed"]] or 0)
        samp = float(r[ix["# Samples"]] or 0)
        if a.op:
            toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
            if not toks or not re.search(a.op, toks[0]):
                continue
        key = (loc[0], loc[1])
        per_line[key][0] += inst
        per_line[key][1] += samp
        for c in stall_cols:
            v = float(r[ix[c]] or 0)
            if v:
                per_line[key][2][c] += v
        tot_i += inst
        tot_s += samp
        sass_rows.append((samp, inst, off, r[ix["Source"]].strip(), key))
    print(f"# total warp instructions {tot_i:.3e}, stall samples {tot_s:.0f}")
    srcs = {}
    for (f, l), (inst, samp, st) in sorted(per_line.items(), key=lambda kv: -(kv[1][0] if a.op else kv[1][1]))[: a.top]:
        if f not in srcs:
            p = os.path.join(os.path.dirname(a.so), "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
        top = ", ".join(f"{k[6:]}:{100 * v / max(samp, 1):.0f}%" for k, v in st.most_common(3))
        print(f"{100 * samp / max(tot_s, 1):5.1f}% samp {100 * inst / max(tot_i, 1):5.1f}% inst  {f}:{l:<4} {text}   [{top}]")
    if a.sass:
        print("# hottest SASS")
        for samp, inst, off, s, key in sorted(sass_rows, reverse=True)[: a.top]:
            print(f"{100 * samp / max(tot_s, 1):5.1f}% samp {100 * inst / max(tot_i, 1):5.2f}% inst  +{off:05x} {key[0]}:{key[1]:<4} {s[:80]}")


if __name__ == "__main__":
    main()
