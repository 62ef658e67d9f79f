#!/bin/bash
# gpurun with retries on "nothing was charged" answers (busy pod / no box): usage tools/gpurun_retry.sh [gpurun options] -- cmd
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -${TAIL:-40}
  if echo "$out" | grep -q "status=transient\|status=busy\|nothing was charged"; then sleep 120; continue; fi
  break
done
