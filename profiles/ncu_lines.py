#!/usr/bin/env python
"""Executed warp-instructions per CUDA-C source line from `ncu --page source --csv --print-source cuda,sass` (gzipped or not).
usage: python profiles/ncu_lines.py X.source.csv[.gz] [top]"""
import csv, gzip, sys, os
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)
cur, hdr, lines = "?", None, []
for r in csv.reader(f):
    if len(r) == 2 and r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; X = r.index("Instructions Executed"); N = r.index("# Samples"); continue
    if hdr is None or not r or not r[0].strip(): continue
    try: lines.append((float(r[X] or 0), float(r[N] or 0), cur, int(r[0]), r[1].strip()[:140]))
    except ValueError: pass
tot = sum(l[0] for l in lines); ts = sum(l[1] for l in lines)
print(f"total warp-instructions: {int(tot)}   samples: {int(ts)}")
byfile = {}
for x, n, c, ln, t in lines: byfile[c] = byfile.get(c, 0) + x
print("per file:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])))
for x, n, c, ln, t in sorted(lines, reverse=True)[:top]:
    print(f"{100*x/tot:6.2f}% inst {100*n/max(ts,1):6.2f}% smp | {c}:{ln}: {t}")
