#!/bin/bash
# round 2, call l: ncu --set full of the single-pass attention backward (predictor shape)
mkdir -p gpurun_out
T=gpurun_out
timeout 100 python tools/ncu_cases.py attn_pred > $T/r2l_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:fa_bwd_umma -o $T/r2l_prof_attn_bwd_sp -f python tools/ncu_cases.py attn_pred > $T/r2l_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2l_ncu.log
