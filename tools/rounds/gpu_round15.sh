#!/bin/bash
# round 15: KV-pass column statistics staged through shared memory; official bench line (with cpu_baseline);
# ncu launch list of one resident step
mkdir -p gpurun_out
T=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k attention > $T/pytest15_attn.log 2>&1
rc=$?; echo "attention rc=$rc"; tail -3 $T/pytest15_attn.log
if [ $rc -ne 0 ]; then echo "attention broken"; exit 1; fi
timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1n.log 2>&1
echo "== attn"; grep -E "fa_" $T/kernel_bench_attn_r1n.log | cut -c1-200
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > $T/pytest15.log 2>&1
echo "pytest all rc=$?"; tail -3 $T/pytest15.log
timeout 420 python bench.py --prof-dump $T/prof_dump_r1n.csv > $T/bench_r1n.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1n.log | cut -c1-2500
python tools/step_breakdown.py $T/prof_dump_r1n.csv 12
timeout 120 python bench.py --steps 1 --warmup 3 --profile-only > $T/plain_r1n.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 7300 -c 2600 --csv --log-file $T/launches_r1n.csv python bench.py --steps 1 --warmup 3 --profile-only > $T/ncu_launch_r1n.log 2>&1
echo "ncu launch list rc=$?"; tail -2 $T/ncu_launch_r1n.log; wc -l $T/launches_r1n.csv
