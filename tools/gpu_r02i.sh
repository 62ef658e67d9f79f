#!/bin/bash
# One-GPU call: whole GPU suite, bench line (with both scoring tables), refine-pass timings per model.
set -u
OUT=gpurun_out/${1:-r02i}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json | cut -c1-400
echo "== tune_mm" | tee $OUT/tune_mm.txt
for mdl in "plane3 10000000" "sphere3 10000000" "circle2 10000000" "line2d 10000000" "absor 1000000" "sphere8 2000000" "plane8 2000000"; do TUNE_MM_ONLY_SHIPPED=1 timeout 200 python tools/tune_mm.py $mdl 2>&1 | tail -1 | tee -a $OUT/tune_mm.txt; done
