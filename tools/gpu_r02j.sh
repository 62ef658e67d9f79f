#!/bin/bash
# Round 2 evidence call (one GPU): ncu --set full captures of the three shipped kernels at the bench configuration and the
# launch list of the bench command.  Numbers under the profiler are evidence for behaviour, never bench values.
set -u
OUT=gpurun_out/${1:-r02j}
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
echo "== ncu: refine pass (mask_moments_kernel, plane3, 10 M points)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_moments -s 3 -c 1 -o $OUT/mask_moments env TUNE_MM_ONLY_SHIPPED=1 python tools/tune_mm.py plane3 10000000 > $OUT/ncu_mm.log 2>&1
echo "rc=$?"
echo "== ncu: refine pass (mask_moments_kernel, sphere3, 10 M points)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_moments -s 3 -c 1 -o $OUT/mask_moments_sphere env TUNE_MM_ONLY_SHIPPED=1 python tools/tune_mm.py sphere3 10000000 > $OUT/ncu_mm_sphere.log 2>&1
echo "rc=$?"
echo "== ncu: consensus_cb at the bench configuration (10 M x 1 M)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 50 -c 1 -o $OUT/consensus_cb $CMD > $OUT/ncu_cb.log 2>&1
echo "rc=$?"
echo "== ncu: fp64 validation kernel (10 M x 16 384)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:consensus_kernel -s 1 -c 1 -o $OUT/consensus_fp64 python bench.py --steps 1 --warmup 1 --precision fp64 --hyps 16384 --no-cpu-baseline --no-e2e --no-configs > $OUT/ncu_fp64.log 2>&1
echo "rc=$?"
# text summaries are made on the box; the reports themselves exceed what gpurun carries back (64 MiB), only the hot kernel's is kept
for r in mask_moments mask_moments_sphere consensus_cb consensus_fp64; do python tools/ncu_summary.py $OUT/$r.ncu-rep > $OUT/$r.txt 2>&1; done
rm -f $OUT/mask_moments.ncu-rep $OUT/mask_moments_sphere.ncu-rep $OUT/consensus_fp64.ncu-rep
echo "== ncu: launch list of bench.py --steps 1"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "rc=$?"
ls -la $OUT
