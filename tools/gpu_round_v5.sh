#!/bin/bash
# end-of-round evidence for the current kernels: ncu launch list of the bench command, compute-sanitizer memcheck
set -u
OUT=gpurun_out/r01_v5
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --points 1000000 --hyps 262144 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches_run.log 2>&1
echo "launch list rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_smoke.py --big > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?"
tail -3 $OUT/memcheck.log
