#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, the bench line, then the evidence the design argues from.
# Usage: tools/gpu_round.sh <tag> [evidence]        (evidence: also ncu launch list + full capture, memcheck, other configs)
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee -a $OUT/pytest.log
echo "== smoke" | tee $OUT/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log
echo "== bench" | tee $OUT/bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>$OUT/bench.err | tail -1 | tee $OUT/bench.json
[ "${2:-}" = "evidence" ] || exit 0
CMD="python bench.py --steps 2 --warmup 1 --points 1000000 --hyps 262144 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv $CMD > $OUT/launches_run.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:consensus_cb -s 40 -c 1 -o $OUT/consensus_cb $CMD > $OUT/full_run.log 2>&1
echo "full capture rc=$?  (summarise here with: python tools/ncu_summary.py $OUT/consensus_cb.ncu-rep)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_smoke.py --big > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -1 $OUT/memcheck.log
timeout 900 python tools/bench_configs.py > $OUT/other_configs.jsonl 2> $OUT/other_configs.err
wc -l $OUT/other_configs.jsonl
