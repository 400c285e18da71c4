#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (stdin): per kernel, the SASS lines with the most stall samples."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, rows=[])
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
seen = set()
for b in blocks:
    if b["name"] in seen:
        continue
    seen.add(b["name"])
    hdr = b["hdr"]
    si = hdr.index("# Samples")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in b["rows"] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in body)
    print(f"== {b['name'][:110]}\n   total samples {tot}, {len(body)} SASS lines")
    agg = {}
    for r in body:
        n = int(r[si])
        for i in st:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
        if n > tot * frac:
            top = {hdr[i][6:]: int(r[i]) for i in st if int(r[i] or 0) > n * 0.2}
            print(f"   {n:7d} {100.0 * n / tot:5.1f}%  {r[1].strip()[:64]:64s} {top}")
    print("   by reason:", [(k[6:], round(100.0 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]])
