#!/bin/bash
# One-GPU call after the fp32 form changes (line2 / line3 / pivot / ultrasound): parity tests, bench line with the scoring table.
set -u
OUT=gpurun_out/${1:-r02g}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json | cut -c1-600
