#!/bin/bash
# R&D: liblsqr_b200 variants differing in the streaming refine kernel's blocking: "<ring KB> <tile KB> <threads>" ...
set -e
cd "$(dirname "$0")/../lsqrrecipes_b200/csrc"
make -s -j6 >/dev/null
mkdir -p ../../tools/bin/variants
for cfg in "$@"; do
  set -- $cfg
  tag="mm_$1_$2_$3"
  F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -DLSQR_MM_SMEM_KB=$1 -DLSQR_MM_TILE_KB=$2 -DLSQR_MM_THREADS=$3"
  nvcc $F -c k_refine.cu -o /tmp/kr_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/variants/lib_$tag.so build/engine.o build/k_score.o build/k_fast.o /tmp/kr_$tag.o build/k_batch.o build/k_bench.o -cudart shared -ldl -lpthread
  echo built $tag
done
