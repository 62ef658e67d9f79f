#!/bin/bash
# Multi-GPU call (gpurun --gpus N): native group context vs one GPU, reference arm and bench under torchrun, compute() on the group.
set -u
N=${2:-8}
OUT=gpurun_out/${1:-r02k}
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
echo "== pytest multi gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -15 | tee -a $OUT/pytest.log
echo "== bench reference arm under torchrun" | tee $OUT/bench_ref.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 3 --warmup 1 --impl reference 2>$OUT/bench_ref.err | tail -1 | tee $OUT/bench_ref.json | cut -c1-300
echo "== bench" | tee $OUT/bench.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus $N --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json | cut -c1-300
echo "== compute probe on the group" | tee $OUT/compute_probe.txt
timeout 300 python tools/compute_probe.py plane3 10000000 0 2>&1 | grep -v "^NCCL" | tee -a $OUT/compute_probe.txt
echo "== dropin demo (multi-GPU case)" | tee $OUT/demo.log
timeout 300 ./examples/dropin_demo 2>&1 | grep -A2 "2 M points" | tee -a $OUT/demo.log
