#!/usr/bin/env python
"""Per-source-line instruction/stall attribution from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur=None; agg={}; nk=0
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0]=='File Path': cur=row[1].split('/')[-1]; continue
    if row[0]=='Function Name': nk+=1; continue
    if row[0]=='Line No': hdr=row; iE=hdr.index('Instructions Executed'); iS=hdr.index('# Samples'); iT=hdr.index('Thread Instructions Executed'); continue
    if row[0]!='':
        try:
            key=(cur,int(row[0]),row[1].strip()[:110]); a=agg.setdefault(key,[0,0,0]); a[0]+=int(row[iE]); a[1]+=int(row[iS]); a[2]+=int(row[iT])
        except Exception: pass
tot=sum(v[0] for v in agg.values()); totS=sum(v[1] for v in agg.values())
print(f"total warp-instructions (all captured launches): {tot}  samples: {totS}")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% inst {100*v[1]/max(totS,1):5.1f}% smp  thr/inst {v[2]/max(v[0],1):4.1f}  {k[0]}:{k[1]}  {k[2]}")
