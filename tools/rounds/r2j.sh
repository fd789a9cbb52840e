#!/bin/bash
# round 2, call j: launch list of the predictor-shaped attention backward (which kernels run, how long)
mkdir -p gpurun_out
T=gpurun_out
timeout 100 python tools/ncu_cases.py attn_pred > $T/r2j_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $T/r2j_launches.csv python tools/ncu_cases.py attn_pred > $T/r2j_ncu.log 2>&1
echo "rc=$?"; grep -E "fa_|copy_rows|delta" $T/r2j_launches.csv | cut -d, -f5,15- | cut -c1-200
