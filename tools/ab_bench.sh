#!/bin/bash
# A/B of library builds: tools/ab_bench.sh A B C ...  (subread_b200/libsubread_cuda_<X>.so); extra env passes through.
mkdir -p gpurun_out
for v in "$@"; do
  SRCUDA_LIB=$PWD/subread_b200/libsubread_cuda_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_$v.log 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{v}.log").read().strip().split("\n")[-1])
    print(f"variant {v}: {d['value']/1e6:.1f} Mreads/s kernel_ms {d['roofline']['kernel_ms']:.2f} e2e {d['e2e']['value']/1e6:.1f} ok={d['e2e']['matches_device_result']} mapped={d['mapped_fraction']}")
except Exception as e:
    print("variant", v, "failed", e); print(open(f"gpurun_out/ab_{v}.log").read()[-600:])
PY
done
