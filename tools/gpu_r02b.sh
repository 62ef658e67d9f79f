#!/bin/bash
# Round 2 GPU call: full parity suite (no -x), refine-kernel variants, compute() latency probe, bench line.
set -u
OUT=gpurun_out/${1:-r02b}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee -a $OUT/pytest.log
echo "== smoke" | tee $OUT/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log
echo "== tune_mm" | tee $OUT/tune_mm.txt
timeout 600 python tools/tune_mm.py plane3 10000000 2>&1 | tee -a $OUT/tune_mm.txt
timeout 300 python tools/tune_mm.py sphere3 10000000 2>&1 | tee -a $OUT/tune_mm.txt
timeout 300 python tools/tune_mm.py absor 1000000 2>&1 | tee -a $OUT/tune_mm.txt
echo "== compute probe" | tee $OUT/compute_probe.txt
timeout 600 python tools/compute_probe.py plane3 10000000 2>&1 | tee -a $OUT/compute_probe.txt
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
