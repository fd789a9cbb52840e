#!/bin/bash
# round 2, call h: ncu --set full (source counters) of the attention forward with the interleaved softmax
mkdir -p gpurun_out
T=gpurun_out
timeout 100 python tools/ncu_cases.py attn_target attn_pred > $T/r2h_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:fa_fwd_umma -o $T/r2h_prof_attn_fwd -f python tools/ncu_cases.py attn_target attn_pred > $T/r2h_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2h_ncu.log
