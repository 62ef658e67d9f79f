#!/bin/bash
# R&D: builds liblsqr_b200 variants that differ in the constant-bank kernel's blocking of ONE model (k_fast.cu cb_blocking override).
#   tools/build_variants.sh "<model id> <R> <PPI>" ...   -> tools/bin/variants/lib_m<id>_<R>x<PPI>.so    (time them with tools/tune_cb.py <model name>)
set -e
cd "$(dirname "$0")/../lsqrrecipes_b200/csrc"
make -s -j6 >/dev/null
mkdir -p ../../tools/bin/variants
for cfg in "$@"; do
  set -- $cfg
  tag="m$1_$2x$3"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr \
    -DLSQR_CB_OVR_MODEL=$1 -DLSQR_CB_OVR_R=$2 -DLSQR_CB_OVR_PPI=$3 -Xptxas -v -c k_fast.cu -o /tmp/kf_$tag.o 2>&1 \
    | grep -A2 "consensus_cb_kernelILi$1E" | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $tag"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/variants/lib_$tag.so build/engine.o build/k_score.o /tmp/kf_$tag.o build/k_refine.o build/k_batch.o build/k_bench.o -cudart shared -ldl -lpthread
done
