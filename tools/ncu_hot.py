#!/usr/bin/env python3
"""Top stall locations of one kernel from an `ncu --page source --csv` dump (SASS view).
    ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
total = sum(int(r[col["# Samples"]] or 0) for r in data if len(r) > col["# Samples"])
print(rows[0][1][:120], " total samples", total)
agg = {n: 0 for n in stall_cols}
for r in data:
    for n in stall_cols:
        agg[n] += int(r[col[n]] or 0)
print("by reason:", ", ".join(f"{n[6:]} {v * 100 // max(total, 1)}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = data[i]
    s = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s * 100.0 / max(total, 1):5.1f}%  {r[col['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
