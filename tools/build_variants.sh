#!/bin/bash
# R&D: builds liblsqr_b200 variants that differ in the constant-bank kernel's compile-time knobs.
#   tools/build_variants.sh "<minblocks> <R_plane> <unused> <threads> [ppi] [unused] [R_all PPI_all]" ...   -> tools/bin/variants/lib_<mb>_<R>_<mix>_<T>_<ppi>_<raw>.so
set -e
cd "$(dirname "$0")/../lsqrrecipes_b200/csrc"
make -s -j4 >/dev/null
for cfg in "$@"; do
  set -- $cfg
  ppi=${5:-2}
  raw=${6:-0}
  sweep=""
  if [ -n "${7:-}" ]; then sweep="-DLSQR_CB_SWEEP_R=$7 -DLSQR_CB_SWEEP_PPI=$8"; fi
  tag="$1_$2_$3_$4_${ppi}_${raw}${7:+_all$7x$8}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr \
    -DLSQR_CB_MINBLOCKS=$1 -DLSQR_CB_R_PLANE=$2 -DLSQR_CB_THREADS=$4 -DLSQR_CB_PPI_PLANE=$ppi $sweep -Xptxas -v -c k_fast.cu -o /tmp/kf_$tag.o 2>&1 \
    | grep -A2 "consensus_cb_kernelILi0E" | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $tag"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/variants/lib_$tag.so build/engine.o build/k_score.o /tmp/kf_$tag.o build/k_refine.o build/k_batch.o build/k_bench.o -cudart shared
done
