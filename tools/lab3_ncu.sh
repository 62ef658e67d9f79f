#!/bin/bash
mkdir -p gpurun_out/lab3_ncu
ncu --set full --import-source on --clock-control none -k regex:cb_kernel --launch-skip 2 -c 1 -o gpurun_out/lab3_ncu/cb_setp_r12 -f ./tools/bin/consensus_lab3 > gpurun_out/lab3_ncu/run.log 2>&1
ncu -i gpurun_out/lab3_ncu/cb_setp_r12.ncu-rep --page raw --csv > gpurun_out/lab3_ncu/cb_setp_r12_raw.csv 2>/dev/null
ls -la gpurun_out/lab3_ncu
