#!/bin/bash
# One-GPU call after widening the template space (34 model ids): the whole GPU suite without -x, the C++ demo.
set -u
OUT=gpurun_out/${1:-r02h}
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -80 | tee -a $OUT/pytest.log
echo "== dropin demo" | tee $OUT/demo.log
timeout 600 ./examples/dropin_demo 2>&1 | tail -40 | tee -a $OUT/demo.log
