#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` export: stall samples per SASS instruction, top N, plus totals per stall reason.
usage: python tools/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("total samples", tot, " instructions", len(body))
per = {h: sum(int(r[ix[h]] or 0) for r in body if len(r) > ix[h]) for h in stall_cols}
print("by reason:", ", ".join(f"{k[6:]}={v} ({100*v/tot:.1f}%)" for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
base = int(body[0][0], 16)
top = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:N]
for r in sorted(top, key=lambda r: int(r[0], 16)):
    s = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols if len(r) > ix[h]), reverse=True)[:2]
    print(f"  +{int(r[0],16)-base:05x} {s:6d} {100*s/tot:5.2f}%  {r[1].strip():60s} {why}")
