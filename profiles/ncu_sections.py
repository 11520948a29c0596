#!/usr/bin/env python
"""Executed SASS instructions per section of a kernel (groups of CUDA-C line ranges), with the opcode mix of each section, from
`ncu --page source --csv --print-source cuda,sass`. usage: python profiles/ncu_sections.py X.source.csv.gz PIXELS sections.py
(sections.py calls grp(name, lambda (file, line): bool)); inlined intrinsics are listed under their own header file too."""
import csv,gzip,os,collections,sys
f=gzip.open(sys.argv[1],'rt')
PX=float(sys.argv[2])
cur='?';line=None;hdr=None
per=collections.defaultdict(lambda: collections.Counter())
tot=0
for r in csv.reader(f):
    if len(r)==2 and r[0]=='File Path': cur=os.path.basename(r[1]); continue
    if len(r)==2: continue
    if r and r[0]=='Line No': hdr=r; X=r.index('Instructions Executed'); continue
    if hdr is None or not r: continue
    if r[0].strip(): line=(cur,int(r[0])); continue
    ins=r[3].strip().split()
    if not ins: continue
    op=ins[0] if not ins[0].startswith('@') else ins[1]
    try: n=float(r[X] or 0)
    except ValueError: continue
    per[line][op]+=n; tot+=n
def grp(name,pred):
    c=collections.Counter()
    for k,v in per.items():
        if pred(k): c.update(v)
    s=sum(c.values())
    print(f"== {name}: {100*s/tot:.1f}%  = {s*32/PX:.1f} instr/px")
    print("   ",", ".join(f"{o} {n*32/PX:.1f}" for o,n in c.most_common(16)))
exec(open(sys.argv[3]).read())
