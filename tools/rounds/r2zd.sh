#!/bin/bash
# round 2, call zd: ncu launch list (gpu__time_duration.sum only, no replay) of ONE step of the final build
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python bench.py --profile-only --steps 1 --warmup 0 > $T/r2zd_plain.log 2>&1 || { echo "plain run failed"; tail -5 $T/r2zd_plain.log; exit 1; }
cat $T/r2zd_plain.log | tail -2 | cut -c1-200
timeout 1100 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $T/r2zd_launches.csv \
  python bench.py --profile-only --steps 1 --warmup 0 > $T/r2zd_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2zd_ncu.log | cut -c1-200; wc -l $T/r2zd_launches.csv
