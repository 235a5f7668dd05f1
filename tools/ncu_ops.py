#!/usr/bin/env python
"""opcode histogram (executed thread-instructions per element) of one kernel in an ncu report:
   python tools/ncu_ops.py report.ncu-rep kernel_regex n_elements"""
import csv, io, re, subprocess, sys, collections
rep, kern, nelem = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
first = re.split(r'(?m)^"Kernel Name",', raw)[1]
rows = list(csv.reader(io.StringIO("\n".join(first.splitlines()[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); tot = 0
for r in rows[1:]:
    if not r or not r[0].startswith("0x"): continue
    toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    if op in ("LDS", "STS", "LDG", "STG", "SHFL", "MUFU", "BAR"): op = ".".join(toks[0].split(".")[:3])
    n = float(r[ix["Instructions Executed"]] or 0)
    cnt[op] += n; tot += n
print(f"total warp instr {tot:.4e} -> {tot * 32 / nelem:.2f} per element")
for op, n in cnt.most_common(40):
    print(f"  {op:24s} {n * 32 / nelem:7.2f}")
