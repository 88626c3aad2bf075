#!/usr/bin/env python
"""tools/ptxas_summary.py [FILTER] -- registers, stack and spills per kernel from subread_b200/csrc/ptxas_sr_api.log (nvcc -Xptxas -v)."""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(root, "subread_b200", "csrc", "ptxas_sr_api.log")).read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for name, stack, ss, sl, regs in pat.findall(log):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void sr::", "")
    if flt in d:
        print(f"{d:48s} regs {regs:>3s}  stack {stack:>4s}  spill st/ld {ss}/{sl}")
