#!/bin/bash
# one --set full capture of the constant-bank consensus kernel on the bench command; numbers under ncu are never bench values
set -u
TAG=${1:-r01_v5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --points 1000000 --hyps 262144 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 40 -c 1 -o $OUT/consensus_cb $CMD > $OUT/full_run.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
