#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep [N] -- executed warp instructions, thread efficiency and stall samples per SOURCE line
(needs -lineinfo and `ncu --import-source on`).  Prints the N heaviest lines per file."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, agg, total = None, None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    src = r[1]
    try:
        line = int(r[0])
        inst = int(d["Instructions Executed"] or 0)
        thr = int(d["Thread Instructions Executed"] or 0)
        smp = int(d["# Samples"] or 0)
    except ValueError:
        continue
    key = (fname, line)
    a = agg.setdefault(key, [0, 0, 0, src])
    a[0] += inst; a[1] += thr; a[2] += smp
    total += inst
tot_s = sum(a[2] for a in agg.values())
print(f"total warp instructions {total}, samples {tot_s}")
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    eff = a[1] / a[0] if a[0] else 0
    print(f"{100*a[0]/max(total,1):5.1f}% inst {100*a[2]/max(tot_s,1):5.1f}% smp  thr/inst {eff:4.1f}  {f}:{line}  {a[3].strip()[:110]}")
