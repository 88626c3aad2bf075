#!/usr/bin/env python
"""tools/ncu_regions.py REPORT -- executed instructions / stall samples of k_map_simple grouped by code region of sr_simple.cuh."""
import subprocess, re, collections, sys
rep = sys.argv[1]
out = subprocess.run(["python", "tools/ncu_lines.py", rep, "2000"], capture_output=True, text=True).stdout.split("\n")
src = open('subread_b200/csrc/sr_simple.cuh').read().split("\n")
def find(pat):
    for i, l in enumerate(src):
        if pat in l:
            return i + 1
marks = [("structs", 1), ("simple_hit", find("bool simple_hit(")), ("pack", find("bool simple_pack")), ("offsets/key", find("int simple_sub_offset")), ("probe", find("int simple_probe")),
         ("general", find("bool simple_general(")), ("primary/indel", find("bool simple_primary_hit(")), ("filter(mode 0/2)", find("unsigned simple_filter(")),
         ("phase: occupancy bits", find("// ---- (1) which units hold anything ----")), ("phase: units", find("// ---- (2) the subreads left: one unit each ----")),
         ("phase: general", find("// ---- (3) everything else, on the general path ----")), ("phase dense", find("bool simple_pass_phase_dense(")),
         ("pass (mode 0/2)", find("bool simple_pass(const DevIndex")),
         ("update_top", find("// update_top_three")), ("merge", find("void simple_merge")), ("kernel", find("k_map_simple(const DevIndex ix"))]
marks = [m for m in marks if m[1]]
agg = collections.OrderedDict((m[0], [0, 0, 0]) for m in marks)
other = [0, 0]
print(out[0])
for l in out[1:]:
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% smp\s+thr/inst\s+([\d.]+)\s+(\S+):(\d+)", l)
    if not m:
        continue
    inst, smp, eff, f, line = float(m[1]), float(m[2]), float(m[3]), m[4], int(m[5])
    text = l.split(f"{f}:{line}")[1].strip()[:25]
    if line - 1 < len(src) and text and text in src[line - 1]:
        name = [mk[0] for mk in marks if mk[1] and mk[1] <= line][-1]
        agg[name][0] += inst; agg[name][1] += smp; agg[name][2] += inst * eff
    else:
        other[0] += inst; other[1] += smp
for k, v in agg.items():
    print(f"{k:14s} inst {v[0]:5.1f}%  smp {v[1]:5.1f}%  thr/inst {v[2]/max(v[0],1e-9):5.1f}")
print("other files    inst %5.1f%%  smp %5.1f%%" % tuple(other))
