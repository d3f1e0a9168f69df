#!/usr/bin/env python
"""Hottest SASS instructions (by warp-stall samples) of one kernel of an ncu report.
    python tools/ncu_hot.py rep.ncu-rep <kernel-index> [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split into per-kernel blocks (each starts with a "Kernel Name" line)
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[kid]
hdr, data = b["rows"][0], [r for r in b["rows"][1:] if len(r) == len(b["rows"][0])]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print(b["name"][:120]); print("instructions", len(data), "samples", tot)
agg = {}
for r in data:
    for c in stall_cols:
        agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
print("stall totals:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} samp {int(r[isamp]):6d} ({100*int(r[isamp])/max(tot,1):4.1f}%) exec {r[iex]:>8s}  {r[isrc].strip()[:80]:80s} {st}")
