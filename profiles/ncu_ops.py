#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples per SASS opcode.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python profiles/ncu_ops.py src.csv"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
S, X, N = ci["Source"], ci["Instructions Executed"], ci["# Samples"]
ex, sm = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= max(S, X, N): continue
    t = r[S].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("I2F", "F2I", "LDG", "LDS", "STS", "STG", "MUFU")) else op.split(".")[0]
    ex[op] += float(r[X] or 0); sm[op] += float(r[N] or 0)
te, ts = sum(ex.values()), sum(sm.values())
print(f"{'opcode':14s} {'inst %':>8s} {'samples %':>10s}")
for k, v in ex.most_common(28):
    print(f"{k:14s} {100*v/te:8.1f} {100*sm[k]/max(ts,1):10.1f}")
print("total warp-instructions executed:", int(te))
