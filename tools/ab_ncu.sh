#!/bin/bash
# tools/ab_ncu.sh VARIANT... -- per variant library (subread_b200/libsubread_cuda_<V>.so): the bench line without the CPU
# baseline, then DRAM bytes / L2 hit rate of one k_map_simple launch (2 M reads) under ncu.  Extra env passes through.
mkdir -p gpurun_out
for v in "$@"; do
  tag=${v}${TAG_SUFFIX}
  SRCUDA_LIB=$PWD/subread_b200/libsubread_cuda_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_$tag.log 2>&1
  SRCUDA_LIB=$PWD/subread_b200/libsubread_cuda_$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum \
     --clock-control none -k regex:k_map_simple -s 2 -c 1 --csv --log-file gpurun_out/abncu_$tag.csv python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python - "$tag" <<'PY'
import json,sys,csv
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{v}.log").read().strip().split("\n")[-1])
    print(f"variant {v}: {d['value']/1e6:.1f} Mreads/s kernel_ms {d['roofline']['kernel_ms']:.3f} e2e {d['e2e']['value']/1e6:.1f} ok={d['e2e']['matches_device_result']} mapped={d['mapped_fraction']} punted={d['roofline']['tier0_punted_fraction']}")
except Exception as e:
    print("variant", v, "failed", e); print(open(f"gpurun_out/ab_{v}.log").read()[-600:])
try:
    rows=[r for r in csv.reader(open(f"gpurun_out/abncu_{v}.csv")) if len(r)>10 and r[0].isdigit()]
    print("   ncu (2M reads):", ", ".join(f"{r[-3].split('__')[-1]}={r[-1]}{r[-2]}" for r in rows))
except Exception as e:
    print("   ncu failed", e)
PY
done
