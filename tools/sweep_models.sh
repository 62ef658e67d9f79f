#!/bin/bash
# R&D: times every library variant under tools/bin/variants on a list of models (fp32 constant-bank kernel)
for m in "$@"; do timeout 200 python tools/tune_cb.py $m 1000000 262144 2>&1 | grep "T evals"; done
