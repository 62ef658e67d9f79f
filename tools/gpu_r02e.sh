#!/bin/bash
set -u
OUT=gpurun_out/${1:-r02e}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== tune_mm" | tee $OUT/tune_mm.txt
for cfg in "plane3 10000000" "plane3 100000" "sphere3 10000000" "absor 1000000" "line2d 10000000" "pivot 1000000" "dense6 1000000" "usxw 200000"; do TUNE_MM_ONLY_SHIPPED=1 timeout 300 python tools/tune_mm.py $cfg 2>&1 | tee -a $OUT/tune_mm.txt; done
echo "== compute probe" | tee $OUT/compute_probe.txt
timeout 600 python tools/compute_probe.py plane3 10000000 2>&1 | tee -a $OUT/compute_probe.txt
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
