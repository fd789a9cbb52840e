#!/bin/bash
# round 18: forward K/V ring of 3 at hd=64 (P in TMEM), split max/sum chains, DACT operand prefetch; ncu of attention
mkdir -p gpurun_out
T=gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k "attention or gemm" > $T/pytest18_ag.log 2>&1
rc=$?; echo "attention+gemm rc=$rc"; tail -3 $T/pytest18_ag.log
if [ $rc -ne 0 ]; then echo "broken"; exit 1; fi
timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1q.log 2>&1
echo "== attn"; grep -E "fa_" $T/kernel_bench_attn_r1q.log | cut -c1-200
timeout 150 python tools/kernel_bench.py gemm > $T/kernel_bench_gemm_r1q.log 2>&1
echo "== gemm"; grep -E "gemm_umma" $T/kernel_bench_gemm_r1q.log | grep -E "dact|fc1" | cut -c1-190
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1q.csv > $T/bench_r1q.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1q.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
timeout 100 python tools/ncu_cases.py attn_pred attn_target > $T/ncu_cases_plain_r1q.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:fa_.*umma -o $T/prof_attn_r1q -f python tools/ncu_cases.py attn_pred attn_target > $T/ncu_attn_r1q.log 2>&1
echo "ncu rc=$?"; tail -2 $T/ncu_attn_r1q.log
