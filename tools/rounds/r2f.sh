#!/bin/bash
# round 2, call f: decoupled (ping-pong) math warpgroups in the attention backward -- numerics, microbench A/B, bench
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "attention" > $T/r2f_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 $T/r2f_pytest.log
if [ $rc -ne 0 ]; then grep -E "rel err|Error|assert" $T/r2f_pytest.log | head -10; fi
timeout 200 python tools/kernel_bench.py attn > $T/r2f_kernel_bench_attn_pp.log 2>&1; grep -E "fa_bwd" $T/r2f_kernel_bench_attn_pp.log | cut -c1-200
AVJ_ATTN_BWD_PP=0 timeout 200 python tools/kernel_bench.py attn > $T/r2f_kernel_bench_attn_nopp.log 2>&1; grep -E "fa_bwd" $T/r2f_kernel_bench_attn_nopp.log | cut -c1-200
if [ $rc -eq 0 ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2f_prof_dump.csv > $T/r2f_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2f_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity']['ok'], d['parity']['grad_rel'])"
fi
