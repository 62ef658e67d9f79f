#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md): per-launch durations of every kernel, then
# one --set full capture of the consensus kernel.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --points 1000000 --hyps 262144 --no-cpu-baseline"   # 184 constant-bank launches + 184 bank refills per step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches_run.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 40 -c 1 -o $OUT/consensus_cb $CMD > $OUT/full_run.log 2>&1
echo "full capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_moments -c 1 -o $OUT/mask_moments python bench.py --steps 1 --warmup 1 --no-cpu-baseline --hyps 65536 > $OUT/full_run2.log 2>&1
echo "refine capture rc=$?"
ls -la $OUT
