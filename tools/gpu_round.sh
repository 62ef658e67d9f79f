#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, a short bench, and ncu evidence.
# Usage: tools/gpu_round.sh <tag> [quick]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== smoke" | tee $OUT/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -5 | tee -a $OUT/bench.log
