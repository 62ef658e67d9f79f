#!/bin/bash
set -u
OUT=gpurun_out/${1:-r02d}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== tune_mm" | tee $OUT/tune_mm.txt
for cfg in "plane3 10000000" "plane3 100000" "plane3 1250000" "sphere3 10000000" "absor 1000000" "line2d 10000000" "pivot 1000000"; do timeout 300 python tools/tune_mm.py $cfg 2>&1 | tee -a $OUT/tune_mm.txt; done
echo "== fp32 band" | tee $OUT/fp32_band.jsonl
timeout 900 python tools/measure_fp32_band.py 200000 20000 2>&1 | tee -a $OUT/fp32_band.jsonl
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
