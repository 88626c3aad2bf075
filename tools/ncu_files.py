#!/usr/bin/env python
"""tools/ncu_files.py REPORT.ncu-rep [N] -- warp instructions, threads per instruction and stall samples per source FILE and for the N
heaviest source lines by instructions and by samples (needs -lineinfo and `ncu --import-source on`)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; f = None; agg = {}; lines = {}
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        f = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        line = int(r[0]); smp = int(d["# Samples"] or 0); inst = int(d["Instructions Executed"] or 0); thr = int(d["Thread Instructions Executed"] or 0)
        lsb = int(d.get("stall_long_sb") or 0)
    except ValueError:
        continue
    a = agg.setdefault(f, [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += smp
    lines[(f, line)] = (inst, thr, smp, lsb, r[1].strip()[:105])
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp instructions {ti}, samples {ts}")
for f, a in agg.items():
    if a[0] * 200 > ti:
        print(f"{f:28s} inst {100*a[0]/ti:5.1f}%  thr/inst {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/ts:5.1f}%")
print("-- by instructions")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[2]/ts:5.1f}% smp  thr/inst {v[1]/max(v[0],1):4.1f}  {f}:{l}  {v[4]}")
print("-- by samples (long-scoreboard share of the line's samples)")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*v[2]/ts:5.1f}% smp  lsb {100*v[3]/max(v[2],1):3.0f}%  {100*v[0]/ti:5.1f}% inst  {f}:{l}  {v[4]}")
