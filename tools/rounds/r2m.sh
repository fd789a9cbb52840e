#!/bin/bash
# round 2, call m: GELU' stored by the forward epilogue (dgrad epilogue = multiply) + single-pass attention backward: full suite, bench
mkdir -p gpurun_out
T=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $T/r2m_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -6 $T/r2m_pytest.log | cut -c1-300
timeout 200 python tools/kernel_bench.py gemm > $T/r2m_kernel_bench_gemm.log 2>&1; grep -E "gelu|dact" $T/r2m_kernel_bench_gemm.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2m_prof_dump.csv > $T/r2m_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2m_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity'])"
python tools/step_breakdown.py $T/r2m_prof_dump.csv 60 > $T/r2m_step_breakdown.txt 2>&1
