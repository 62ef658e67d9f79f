#!/bin/bash
# Round 2 GPU call: parity suite, ncu evidence for the shipped kernels at the bench configuration, fp32 band measurement.
set -u
OUT=gpurun_out/${1:-r02c}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee -a $OUT/pytest.log
echo "== fp32 band" | tee $OUT/fp32_band.jsonl
timeout 900 python tools/measure_fp32_band.py 200000 20000 2>&1 | tee -a $OUT/fp32_band.jsonl
echo "== ncu: refine pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_moments -c 1 -o $OUT/mask_moments env TUNE_MM_ONLY_SHIPPED=1 python tools/tune_mm.py plane3 10000000 > $OUT/ncu_mm.log 2>&1
echo "rc=$?"
echo "== ncu: consensus_cb at the bench configuration (10 M x 1 M)"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 50 -c 1 -o $OUT/consensus_cb $CMD > $OUT/ncu_cb.log 2>&1
echo "rc=$?"
echo "== ncu: launch list of bench.py --steps 1"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "rc=$?"
echo "== ncu: fp64 validation kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:consensus_kernel -s 2 -c 1 -o $OUT/consensus_fp64 python bench.py --steps 1 --warmup 1 --precision fp64 --hyps 65536 --no-cpu-baseline --no-e2e --no-configs > $OUT/ncu_fp64.log 2>&1
echo "rc=$?"
ls -la $OUT
