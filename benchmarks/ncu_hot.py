#!/usr/bin/env python
"""Top stalled SASS instructions of every launch in an ncu report (source page), with their dominant stall reasons.
   python benchmarks/ncu_hot.py gpurun_out/prof_x.ncu-rep [topN]"""
import csv, io, subprocess, sys

rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": next(csv.reader([line]))[1], "lines": []}
        blocks.append(cur)
    elif cur is not None:
        cur["lines"].append(line)
for bi, b in enumerate(blocks):
    rows = list(csv.DictReader(io.StringIO("\n".join(b["lines"]))))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"== launch {bi}: {b['name'][:90]}  samples={tot}")
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:topn]
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    for i in sorted(idx):
        r = rows[i]
        n = int(r["# Samples"] or 0)
        st = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"  {100.0 * n / max(tot, 1):5.1f}%  [{i:5d}] {r['Source'].strip()[:70]:70s}  " + ", ".join(f"{k}={v}" for v, k in st if v))
