#!/bin/bash
# one-off A/B of tier-0 variants (round 2): bench line + instruction/issue metrics of one 1 Mi-read launch
mkdir -p gpurun_out
for v in "$@"; do
  SRCUDA_LIB=$PWD/subread_b200/libsubread_cuda_$v.so python tools/old_bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  SRCUDA_LIB=$PWD/subread_b200/libsubread_cuda_$v.so ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers \
     --clock-control none -k regex:k_map_simple -s 2 -c 1 --csv --log-file gpurun_out/abncu_$v.csv python tools/old_bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python - "$v" <<'PY'
import json,sys,csv
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{v}.log").read().strip().split("\n")[-1])
    print(f"variant {v}: {d['value']/1e6:.1f} Mreads/s kernel_ms {d['roofline']['kernel_ms']:.3f} ok={d['e2e']['matches_device_result']} mapped={d['mapped_fraction']}")
except Exception as e:
    print("variant", v, "failed", e); print(open(f"gpurun_out/ab_{v}.log").read()[-600:])
try:
    rows=[r for r in csv.reader(open(f"gpurun_out/abncu_{v}.csv")) if len(r)>10 and r[0].isdigit()]
    print("   ncu:", ", ".join(f"{r[-3].split('__')[-1][:28]}={r[-1]}" for r in rows))
except Exception as e:
    print("   ncu failed", e)
PY
done
