#!/usr/bin/env python
"""tools/ncu_dense_regions.py REPORT.ncu-rep -- warp instructions / threads per instruction / stall samples of the paired-end dense tier by
region of sr_simple.cuh + sr_simple_pe.cuh (regions found by text markers in the CURRENT sources: use on a report of the current build)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; f = None; lines = {}
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): f = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: line = int(r[0]); smp = int(d["# Samples"] or 0); inst = int(d["Instructions Executed"] or 0); thr = int(d["Thread Instructions Executed"] or 0)
    except ValueError: continue
    lines[(f, line)] = (inst, thr, smp)
ti = sum(v[0] for v in lines.values()); ts = sum(v[2] for v in lines.values())
def regions(fname, marks):
    src = open('subread_b200/csrc/' + fname).read().split('\n')
    def find(p):
        for i, l in enumerate(src):
            if p in l: return i + 1
    ms = [(n, find(p)) for n, p in marks]
    ms = sorted([m for m in ms if m[1]], key=lambda m: m[1])
    for i, (n, lo) in enumerate(ms):
        hi = ms[i + 1][1] - 1 if i + 1 < len(ms) else 10 ** 6
        a = b = c = 0
        for (ff, l), v in lines.items():
            if ff == fname and lo <= l <= hi: a += v[0]; b += v[1]; c += v[2]
        if a: print(f"{fname:18s} {n:34s} inst {100*a/ti:5.1f}%  thr/inst {b/max(a,1):5.1f}  smp {100*c/ts:5.1f}%   [{lo}-{hi}]")
print(f"total warp instructions {ti}")
regions("sr_simple.cuh", [("cluster accessors", "struct SimpleCluster {"), ("simple_hit (list walk)", "// A fresh cluster for the hit"), ("pack", "// Pack the thread's read from the staging"),
    ("offsets / keys", "// Offset of the first lookup of subread s"), ("(sparse-table code)", "// What one survivor's bucket record says"),
    ("dense: setup", "bool simple_pass_phase_dense("), ("dense: hit()", "auto hit = [&](uint32_t kv, int l, int u, int sub) -> bool {"),
    ("dense: batch setup, issue", "		const int n_sub = min(kSubs, applied - s_base), base = s_base * GAP;\n"), ("dense: unit loop head", "			const Record8 r = rn;"),
    ("dense: look (stage 2)", "		auto look = [&]"), ("dense: apply (stage 3)", "		auto apply = [&]"), ("dense: pipelined loop", "		issue(0);"),
    ("dense: read entries", "		if (ok && ne) {"), ("dense: slow lookups + runs", "		// the keys with several positions: one flat loop"), ("rest of sr_simple.cuh", "// Asks for the dense units of subreads")])
regions("sr_simple_pe.cuh", [("pe merge", "__device__ __forceinline__ bool simple_pe_merge("), ("pe kernel: stage + pack", "k_map_simple_pe(const DevIndex ix"), ("pe kernel: passes", "	// ---- 2. four passes"), ("pe kernel: output", "	// ---- 3. the warp's bigtable slice")])
regions("sr_common.cuh", [("sr_common.cuh", "#pragma once")])
