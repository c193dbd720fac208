#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counters by CUDA source line / line range.
usage: ncu_lines.py report.ncu-rep [kernel-regex] [--regions a:b:name,...] [--top N]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
top = 40
regions = []
for a in sys.argv[2:]:
    if a.startswith("--top="): top = int(a[6:])
    if a.startswith("--regions="):
        for r in a[10:].split(","):
            lo, hi, nm = r.split(":"); regions.append((nm, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; ie = hdr.index("Instructions Executed"); sa = hdr.index("# Samples")
agg = {}
for r in rows[hi + 1:]:
    if r and r[0].strip().isdigit():
        try:
            a = agg.setdefault(int(r[0]), [r[1].strip(), 0, 0]); a[1] += int(r[ie]); a[2] += int(r[sa])
        except ValueError: pass
tot = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot}  samples {ts}")
for nm, lo, hi_ in regions:
    i = sum(v[1] for k, v in agg.items() if lo <= k <= hi_); s = sum(v[2] for k, v in agg.items() if lo <= k <= hi_)
    print(f"{nm:24s} {100*i/tot:5.1f}% inst ({i/1e6:9.1f} M)  {100*s/ts:5.1f}% samples")
for ln, (src, i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:5d} {100*i/tot:5.1f}% inst {100*s/ts:5.1f}% samp | {src[:100]}")
