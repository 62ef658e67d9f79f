#!/bin/bash
# Multi-GPU call (gpurun --gpus N): native NCCL context, torchrun check in native and hook modes, the C++ drop-in on every GPU,
# the bench at N ranks, compute() latency on the single-process group.
set -u
N=${2:-2}
OUT=gpurun_out/${1:-r02f}
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
echo "== pytest multi gpu" | tee $OUT/pytest.log
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_cxx_dropin.py -m gpu -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== check_multi_gpu (torchrun)" | tee $OUT/check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 tools/check_multi_gpu.py 2>&1 | grep -v "^W\|^\[W\|NCCL version" | tail -40 | tee -a $OUT/check.log
echo "== dropin demo" | tee $OUT/demo.log
timeout 600 ./examples/dropin_demo 2>&1 | tail -12 | tee -a $OUT/demo.log
echo "== compute probe on the group" | tee $OUT/compute_probe.txt
timeout 600 python tools/compute_probe.py plane3 10000000 0 2>&1 | tee -a $OUT/compute_probe.txt
echo "== bench reference arm under torchrun" | tee $OUT/bench_ref.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 3 --warmup 1 --impl reference 2>$OUT/bench_ref.err | tail -1 | tee $OUT/bench_ref.json
echo "== bench" | tee $OUT/bench.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus $N --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29584 bench.py --gpus $N --steps 2 --warmup 3 --comm hooks --no-cpu-baseline 2>$OUT/bench_hooks.err | tail -1 | tee $OUT/bench_hooks.json
