#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep -- the handful of ncu metrics the profiles/ notes quote, per kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"][:60], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"   {k:72s} {d[k]:>18s} {units[hdr.index(k)]}")
    st = []
    for k in hdr:
        if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct") and d[k]:
            try:
                st.append((float(d[k]), k.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:7]:
        print(f"   stall {k:40s} {v:8.1f} % of active warps")
