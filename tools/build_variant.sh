#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=V ...] -- builds subread_b200/libsubread_cuda_NAME.so for A/B runs (tools/ab_bench.sh)
set -e
name=$1; shift
cd "$(dirname "$0")/../subread_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c sr_api.cu -o /tmp/sr_api_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsubread_cuda_$name.so /tmp/sr_api_$name.o sr_index.o sr_fastq.o sr_reader.o -cudart static -Xcompiler -pthread -lz
echo built libsubread_cuda_$name.so
