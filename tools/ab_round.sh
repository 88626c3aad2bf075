#!/bin/bash
# tools/ab_round.sh -- one gpurun call: parity of the dense-table tiers for the shipped library, then A/B bench lines of library variants
# (tools/build_variant.sh) on the human-sized workloads.  Usage: tools/ab_round.sh TAG "variants..." ["ENV=V variants ..."]
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_human.py -m gpu -q -x -k "dense or human or golden" > gpurun_out/pytest_dense_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dense_$TAG.log
tools/ab_bench.sh human_pe $1
shift
for spec in "$@"; do
  envs=${spec%% *}; vars=${spec#* }
  echo "-- with $envs"
  env $envs tools/ab_bench.sh human_pe $vars
done
