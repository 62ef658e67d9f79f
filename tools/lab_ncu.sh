#!/bin/bash
# Per-pipe instruction counts of the lab variants (which pipe does each opcode use?)
mkdir -p gpurun_out/lab_ncu
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum,sm__inst_executed.sum,sm__pipe_alu_cycles_active.sum,sm__pipe_fma_cycles_active.sum,sm__pipe_fmaheavy_cycles_active.sum,sm__pipe_fmalite_cycles_active.sum,sm__cycles_active.sum,smsp__issue_active.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:lab_kernel --csv --log-file gpurun_out/lab_ncu/lab_pipes.csv ./tools/bin/consensus_lab > gpurun_out/lab_ncu/run.log 2>&1
ncu --metrics $M --clock-control none -k regex:pipe_kernel --csv --log-file gpurun_out/lab_ncu/pipe_pipes.csv ./tools/bin/pipe_lab > gpurun_out/lab_ncu/run2.log 2>&1
ls -la gpurun_out/lab_ncu
