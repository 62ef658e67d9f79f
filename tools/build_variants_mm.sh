#!/bin/bash
# R&D: liblsqr_b200 variants differing in the streaming refine kernel's blocking: "<ctas_per_sm> <rows_in_flight>" ...
set -e
cd "$(dirname "$0")/../lsqrrecipes_b200/csrc"
make -s -j4 >/dev/null
for cfg in "$@"; do
  set -- $cfg
  tag="mm_$1_$2"
  F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -DLSQR_MM_CTAS=$1 -DLSQR_MM_U=$2"
  nvcc $F -c k_refine.cu -o /tmp/kr_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/variants/lib_$tag.so build/engine.o build/k_score.o build/k_fast.o /tmp/kr_$tag.o build/k_bench.o -cudart shared
  echo built $tag
done
