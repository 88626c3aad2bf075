run() { wl=$1; tag=$2; shift 2; env "$@" python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_${wl}_$tag.log 2>&1; python - "$wl" "$tag" <<'PY'
import json,sys
wl,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/e2e_{wl}_{t}.log").read().strip().split("\n")[-1]); e=d['e2e']
    print(f"{wl} / {t}: e2e {e['value']/1e6:.1f} Mreads/s  {e['ms_per_step_wall']:.1f} ms/step (kernels {d['ms_per_step']:.1f} ms)  d2h {e['d2h_bytes_per_step']/1e6:.0f} MB ok={e['matches_device_result']}")
except Exception as ex:
    print(wl, t, "failed", ex); print(open(f"gpurun_out/e2e_{wl}_{t}.log").read()[-500:])
PY
}
for wl in cfg2_se human_pe; do
  run $wl pool_default A=1
  run $wl pool_b512k SRCUDA_BATCH_FRAGS=524288
  run $wl pool_b256k SRCUDA_BATCH_FRAGS=262144
  run $wl pool_b128k SRCUDA_BATCH_FRAGS=131072
done
