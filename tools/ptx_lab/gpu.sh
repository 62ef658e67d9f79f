#!/bin/bash
# on the GPU box: time every variant, then per-pipe instruction counts for a few
mkdir -p gpurun_out/ptx_lab
./tools/bin/ptx_lab_run tools/ptx_lab/out | tee gpurun_out/ptx_lab/times.txt
