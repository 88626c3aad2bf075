#!/usr/bin/env python
"""tools/sass_lines.py OBJECT KERNEL_SUBSTRING [N] -- static SASS instruction count per source line of one kernel (needs -lineinfo)."""
import collections, re, subprocess, sys, tempfile, os
obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cnt, fn, cur, total = collections.Counter(), None, None, 0
for line in dis.split("\n"):
    m = re.search(r"\.section\s+\.text\.(\S+?),", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if fn and pat in fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        cnt[cur] += 1
        total += 1
print("total SASS instructions:", total)
src = {}
for (f, l), c in cnt.most_common(top):
    if f not in src:
        for root in ("subread_b200/csrc", "."):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                src[f] = open(pth).read().split("\n")
                break
        else:
            src[f] = []
    text = src[f][l - 1].strip()[:100] if l - 1 < len(src[f]) else ""
    print(f"{c:5d}  {f}:{l}  {text}")
