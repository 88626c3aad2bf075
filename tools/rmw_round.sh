#!/bin/bash
# tools/rmw_round.sh TAG -- what a later block of a split index costs: the chunk merged into a non-zero bigtable, with the
# read-modify-write form of the thread-per-fragment tiers and without it (SRCUDA_TIER_RMW=0: warp kernel).  Run under gpurun.
TAG=${1:-r01}
O=gpurun_out/rmw_$TAG.jsonl
: > $O
run() { timeout "$1" python tools/scale_bench.py "${@:2}" >> $O 2>> gpurun_out/rmw_$TAG.err; echo "rc=$? ${*:2}"; }
run 100 --genome-mbp 100 --modes se,pe --reads 2000000 --second-pass --check 10000
SRCUDA_TIER_RMW=0 run 100 --genome-mbp 100 --modes se,pe --reads 2000000 --second-pass --check 0
run 150 --genome-mbp 3100 --modes se,pe --reads 1000000 --second-pass --check 2000
grep -c . $O
