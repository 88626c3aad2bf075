#!/usr/bin/env python
"""tools/bench_line.py LOG [LABEL] -- one-line digest of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
r = d.get("roofline") or {}
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "%.1f Mreads/s" % (d["value"] / 1e6), "kernel_ms %.2f" % r.get("kernel_ms", 0),
      "tier0_ms %.2f" % r.get("tier0_ms", 0), "punted %.4f" % r.get("tier0_punted_fraction", 0), "frac %.3f" % r.get("frac", 0),
      "e2e %.1f" % (d["e2e"]["value"] / 1e6), "ok=%s" % d["e2e"]["matches_device_result"], "mapped=%s" % d["mapped_fraction"])
