#!/bin/bash
# tools/gpu_round.sh TAG -- the evidence set of one kernel version on a B200 box (run under gpurun): GPU parity tests, the bench
# workloads (both arms on the default one), ncu launch list and ncu --set full captures of the fast tier's kernels.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu_$TAG.txt; nproc >> gpurun_out/gpu_$TAG.txt; free -g | head -2 >> gpurun_out/gpu_$TAG.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_cfg2_se_$TAG.log 2> gpurun_out/bench_cfg2_se_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg2_se_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_cfg2_se_$TAG.log 2>&1; echo "ref rc=$?"
for wl in mixed human_pe human_subjunc cfg2_pe sweep; do
  python bench.py --workload $wl --no-cpu-baseline --steps 3 > gpurun_out/bench_${wl}_$TAG.log 2> gpurun_out/bench_${wl}_$TAG.err; echo "$wl rc=$?"; tail -c 200 gpurun_out/bench_${wl}_$TAG.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_map_simple -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_cfg2_se \
    python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_cfg2_se_$TAG.log 2>&1; echo "ncu full cfg2_se rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_map_simple_pe -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_human_pe \
    python bench.py --workload human_pe --reads 1000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_human_pe_$TAG.log 2>&1; echo "ncu full human_pe rc=$?"
python - "$TAG" <<'PY'
import json,sys,glob
tag=sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/bench_*_{tag}.log")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1]); e=d.get('e2e') or {}
        print(f"{f}: {d['value']/1e6:.1f} M/s e2e {e.get('value',0)/1e6:.1f} M/s")
    except Exception as ex: print(f, "unreadable", ex)
PY
