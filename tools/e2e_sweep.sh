#!/bin/bash
# tools/e2e_sweep.sh -- the host-buffer entry point (bench.py's e2e leg) under different batch sizes / expand threads / plain copy
mkdir -p gpurun_out
run() {  # workload tag env...
  wl=$1; tag=$2; shift 2
  env "$@" python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_${wl}_$tag.log 2>&1
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/e2e_{wl}_{t}.log").read().strip().split("\n")[-1]); e=d['e2e']
    print(f"{wl} / {t}: e2e {e['value']/1e6:.1f} Mreads/s  {e['ms_per_step_wall']:.1f} ms/step (kernels {d['ms_per_step']:.1f} ms)  d2h {e['d2h_bytes_per_step']/1e6:.0f} MB ok={e['matches_device_result']}")
except Exception as ex:
    print(wl, t, "failed", ex); print(open(f"gpurun_out/e2e_{wl}_{t}.log").read()[-500:])
PY
}
for wl in "$@"; do
  run $wl default A=1
  run $wl b256k SRCUDA_BATCH_FRAGS=262144
  run $wl b512k SRCUDA_BATCH_FRAGS=524288
  run $wl b2m SRCUDA_BATCH_FRAGS=2097152
  run $wl t8 SRCUDA_EXPAND_THREADS=8
  run $wl t4 SRCUDA_EXPAND_THREADS=4
  run $wl plain SRCUDA_COMPACT_D2H=0
  run $wl plain_b256k SRCUDA_COMPACT_D2H=0 SRCUDA_BATCH_FRAGS=262144
done
