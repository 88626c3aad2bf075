#!/usr/bin/env python
"""tools/ncu_traffic.py WORKLOAD REPORT.ncu-rep FRAGMENTS_PER_LAUNCH -- records the DRAM traffic of the captured tier-0 launch in
profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, per fragment), stamped with
the hash of the kernel sources it was taken from.  bench.py quotes it as roofline.traffic only while that hash still matches."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

workload, rep, frags = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, row = rows[0], rows[1], rows[2]
d = dict(zip(hdr, row))
u = dict(zip(hdr, units))


def to_bytes(key):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
    return float(d[key]) * scale


total = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    t = json.load(open(path))
except Exception:
    t = {}
h = kernel_source_hash()
if t.get("kernel_source_hash") != h:
    t = {"kernel_source_hash": h, "workloads": {}}
t["workloads"][workload] = {"kernel": d["Kernel Name"], "fragments_per_launch": frags, "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
                            "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "dram_bytes_per_fragment": total / frags,
                            "source": f"ncu --set full --clock-control none, one launch of {frags} fragments ({os.path.basename(rep)})"}
json.dump(t, open(path, "w"), indent=1)
print(json.dumps(t["workloads"][workload]))
