#!/usr/bin/env python
"""Per-phase split of one kernel's ncu source page (SASS view): instructions executed, stall samples and the top stall
reasons between consecutive BAR.SYNCs (= the phases of K1).  usage: ncu -i rep --page source --csv | ncu_phase_split.py"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
seg, segs = 0, defaultdict(lambda: defaultdict(float))
kinds = defaultdict(lambda: defaultdict(float))
for r in rows:
    if not r or not r[0].startswith("0x"):
        continue
    src = r[ix["Source"]].strip()
    s = segs[seg]
    s["inst"] += float(r[ix["Instructions Executed"]] or 0)
    s["samples"] += float(r[ix["# Samples"]] or 0)
    for n in stalls:
        s[n] += float(r[ix[n]] or 0)
    op = src.split()[0].split(".")[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1].split(".")[0]
    kinds[seg][op] += float(r[ix["Instructions Executed"]] or 0)
    if "BAR.SYNC" in src:
        seg += 1
ti = sum(s["inst"] for s in segs.values())
ts = sum(s["samples"] for s in segs.values())
print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
for k in sorted(segs):
    s = segs[k]
    top = sorted(((s[n], n) for n in stalls), reverse=True)[:3]
    ops = sorted(((v, o) for o, v in kinds[k].items()), reverse=True)[:5]
    print(f"seg {k:2d}: inst {s['inst'] / ti * 100:5.1f}%  samples {s['samples'] / max(ts, 1) * 100:5.1f}%  "
          + ", ".join(f"{n[6:]} {v / max(s['samples'], 1) * 100:.0f}%" for v, n in top)
          + "   ops: " + ", ".join(f"{o} {v / max(s['inst'], 1) * 100:.0f}%" for v, o in ops))
