#!/bin/bash
# A/B of library builds: tools/ab_bench.sh WORKLOAD A B C ...  (subread_b200/libsubread_cuda_<X>.so; "default" = the shipped library); extra env passes through.
mkdir -p gpurun_out
wl=$1; shift
for v in "$@"; do
  lib=$PWD/subread_b200/libsubread_cuda_$v.so
  [ "$v" = default ] && lib=$PWD/subread_b200/libsubread_cuda.so
  SRCUDA_LIB=$lib python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ab_${wl}_$v.log 2>&1
  python - "$wl" "$v" <<'PY'
import json,sys
wl,v=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/ab_{wl}_{v}.log").read().strip().split("\n")[-1])
    r=d['roofline']
    print(f"{wl} / {v}: {d['value']/1e6:.1f} Mreads/s  kernel_ms {r['kernel_ms']:.2f} tier0_ms {r['tier0_ms']:.2f} warp_ms {r['warp_kernel_ms']:.2f} scalar_ms {r['scalar_kernel_ms']:.2f} queued {r['tier0_punted_fraction']:.4f} slot0 {d['slot0_is_true_position']}")
except Exception as e:
    print(wl, v, "failed", e); print(open(f"gpurun_out/ab_{wl}_{v}.log").read()[-800:])
PY
done
