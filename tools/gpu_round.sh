#!/bin/bash
# tools/gpu_round.sh TAG -- the evidence set of one kernel version on a B200 box (run under gpurun):
# GPU parity tests, default bench (both arms), ncu launch list and one ncu --set full capture of the voting kernels.
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc > gpurun_out/host.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; tail -c 900 gpurun_out/bench_ref_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_map_ -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
