#!/bin/bash
# tools/scale_round.sh TAG -- the other BASELINE shapes through tools/scale_bench.py on one B200 (run under gpurun).
TAG=${1:-r01}
O=gpurun_out/scale_$TAG.jsonl
: > $O
free -g | head -2 > gpurun_out/scale_host_$TAG.txt; nproc >> gpurun_out/scale_host_$TAG.txt
run() { timeout "$1" python tools/scale_bench.py "${@:2}" >> $O 2>> gpurun_out/scale_$TAG.err; echo "rc=$? ${*:2}"; }
run 400 --genome-mbp 100 --sweep 50,100,150,250 --reads 2000000
run 400 --genome-mbp 100 --modes pe --sweep 100,150 --reads 2000000
run 400 --genome-mbp 100 --modes sj-se,sj-pe --read-len 100 --reads 200000
run 400 --genome-mbp 100 --full --modes se,pe --reads 1000000
run 1200 --genome-mbp 3100 --modes se,pe --reads 1000000 --check 10000
python - "$O" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    try: d=json.loads(l)
    except Exception: continue
    print(f"{d['workload'][:110]:110s} {d['reads_per_s']/1e6:9.1f} Mreads/s ms {d['ms_per_step']:8.2f} punt {d['tier0_punted_fraction']:.3f} ovf {d['overflow_fraction']:.4f} true {d['slot0_is_true_position']:.4f} oracle {d.get('oracle_equal')} {d.get('oracle_junc_equal','')} build {d['index_build_s']:.0f}s")
PY
tail -5 gpurun_out/scale_$TAG.err
