#!/bin/bash
# Runs on the GPU box (via gpurun).  Usage: tools/gpu_round.sh <tag> <mode> [N]
#   check      GPU parity tests, smoke, the bench line, single-launch refine timings per estimator      (one GPU)
#   evidence   ncu --set full captures of the shipped kernels at the bench configuration + launch list  (one GPU)
#   sanitize   compute-sanitizer memcheck over every model and entry point (tools/sanitize_smoke.py)     (one GPU)
#   racecheck  compute-sanitizer racecheck over a subset of the models (shared-memory hazards)             (one GPU)
#   multi N    multi-GPU tests, reference arm and bench under torchrun, compute() on the group, C++ demo (gpurun --gpus N)
# Everything lands in gpurun_out/<tag>/ (text only: ncu reports are summarised on the box, gpurun carries back <= 64 MiB).
set -u
TAG=${1:-r02}
MODE=${2:-check}
N=${3:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
case $MODE in
check)
  echo "== pytest -m gpu" | tee $OUT/pytest.log
  timeout 1500 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -50 | tee -a $OUT/pytest.log
  echo "== smoke" | tee $OUT/smoke.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log
  echo "== bench" | tee $OUT/bench.log
  timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json | cut -c1-400
  echo "== refine pass per estimator" | tee $OUT/tune_mm.txt
  for mdl in "plane3 10000000" "sphere3 10000000" "circle2 10000000" "line2d 10000000" "absor 1000000" "sphere8 2000000" "plane8 2000000"; do
    TUNE_MM_ONLY_SHIPPED=1 timeout 200 python tools/tune_mm.py $mdl 2>&1 | tail -1 | tee -a $OUT/tune_mm.txt
  done
  ;;
evidence)
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs"
  NCU="ncu --set full --clock-control none --import-source on"
  timeout 600 $NCU -k regex:mask_moments -s 3 -c 1 -o $OUT/mask_moments env TUNE_MM_ONLY_SHIPPED=1 python tools/tune_mm.py plane3 10000000 > $OUT/ncu_mm.log 2>&1; echo "refine pass, plane rc=$?"
  timeout 600 $NCU -k regex:mask_moments -s 3 -c 1 -o $OUT/mask_moments_sphere env TUNE_MM_ONLY_SHIPPED=1 python tools/tune_mm.py sphere3 10000000 > $OUT/ncu_mm_sphere.log 2>&1; echo "refine pass, sphere rc=$?"
  timeout 900 $NCU -k regex:consensus_cb -s 50 -c 1 -o $OUT/consensus_cb $CMD > $OUT/ncu_cb.log 2>&1; echo "consensus_cb at 10 M x 1 M rc=$?"
  timeout 600 $NCU -k regex:consensus_kernel -s 1 -c 1 -o $OUT/consensus_fp64 python bench.py --steps 1 --warmup 1 --precision fp64 --hyps 16384 --no-cpu-baseline --no-e2e --no-configs > $OUT/ncu_fp64.log 2>&1; echo "fp64 kernel rc=$?"
  for r in mask_moments mask_moments_sphere consensus_cb consensus_fp64; do python tools/ncu_summary.py $OUT/$r.ncu-rep > $OUT/$r.txt 2>&1; done
  rm -f $OUT/mask_moments.ncu-rep $OUT/mask_moments_sphere.ncu-rep $OUT/consensus_fp64.ncu-rep
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
  ;;
sanitize)
  timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_smoke.py --big > $OUT/memcheck.log 2>&1
  echo "memcheck rc=$?"; tail -8 $OUT/memcheck.log
  ;;
profile_cb)   # ncu --set full of the constant-bank kernel for the models named in $CB_MODELS (default: the ones furthest from the pipe)
  for mdl in ${CB_MODELS:-uscp usxw ray circle2}; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 3 -c 1 -o $OUT/cb_$mdl python tools/tune_cb.py $mdl 1000000 262144 > $OUT/ncu_cb_$mdl.log 2>&1; echo "$mdl rc=$?"
    python tools/ncu_summary.py $OUT/cb_$mdl.ncu-rep > $OUT/cb_$mdl.txt 2>&1
    ncu -i $OUT/cb_$mdl.ncu-rep --page source --csv 2>/dev/null | head -400 > $OUT/cb_${mdl}_source.csv
    rm -f $OUT/cb_$mdl.ncu-rep
  done
  ;;
run)          # an arbitrary command: CMD="python tools/..."
  timeout 900 bash -c "$CMD" 2>&1 | tail -60 | tee $OUT/run.log
  ;;
only)         # a subset of the GPU tests: K="<pytest -k expression>"
  timeout 900 python -m pytest tests -m gpu -q -k "${K:-degenerate}" 2>&1 | tail -40 | tee $OUT/pytest_only.log
  ;;
batch)        # the batched path only: its GPU tests and the batch probe
  timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_full_size_gpu.py -m gpu -q -k "batch" 2>&1 | tail -30 | tee $OUT/pytest_batch.log
  timeout 300 python tools/batch_probe.py 2>&1 | tail -20 | tee $OUT/batch_probe.txt
  ;;
sweep_cb)     # times the blocking variants of tools/build_variants.sh against the shipped library, per model in $CB_MODELS
  for mdl in ${CB_MODELS:-uscp usxw ray circle2 sphere3}; do timeout 300 python tools/tune_cb.py $mdl 1000000 262144 2>&1 | grep "T evals" | tee -a $OUT/sweep_cb.txt; done
  ;;
racecheck)
  SMOKE_MODELS="plane3 sphere3 line2d absor uscp sphere8 dense5" timeout 1700 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_smoke.py > $OUT/racecheck.log 2>&1
  echo "racecheck rc=$?"; grep -c "hazard" $OUT/racecheck.log; tail -5 $OUT/racecheck.log
  ;;
multi)
  echo "== pytest multi gpu" | tee $OUT/pytest.log
  timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_cxx_dropin.py -m gpu -q 2>&1 | tail -15 | tee -a $OUT/pytest.log
  echo "== bench reference arm under torchrun" | tee $OUT/bench_ref.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 3 --warmup 1 --impl reference 2>$OUT/bench_ref.err | tail -1 | tee $OUT/bench_ref.json | cut -c1-300
  echo "== bench" | tee $OUT/bench.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus $N --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json | cut -c1-300
  echo "== compute probe on the group" | tee $OUT/compute_probe.txt
  timeout 300 python tools/compute_probe.py plane3 10000000 0 2>&1 | grep -v "^NCCL" | tee -a $OUT/compute_probe.txt
  echo "== dropin demo (multi-GPU case)" | tee $OUT/demo.log
  timeout 300 ./examples/dropin_demo 2>&1 | grep -A2 "2 M points" | tee -a $OUT/demo.log
  ;;
*) echo "unknown mode $MODE"; exit 2 ;;
esac
