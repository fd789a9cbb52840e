#!/bin/bash
# round 2, call d: MUFU / pack / FMA-poly throughput microbenchmark (settles the attention softmax denominator), plain and
# with ncu pipe counters
mkdir -p gpurun_out
T=gpurun_out
./tools/micro/mufu_bw > $T/r2d_mufu_bw.txt 2>&1 && \
ncu --metrics sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,sm__cycles_elapsed.max,sm__cycles_active.avg,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file $T/r2d_mufu_ncu.csv ./tools/micro/mufu_bw > $T/r2d_mufu_ncu_stdout.txt 2>&1
echo "rc=$?"; cat $T/r2d_mufu_bw.txt
