#!/bin/bash
# round 2, call g: group-interleaved softmax (value-neutral dependency chain) in the attention forward -- numerics, microbench A/B
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -x -k "attention" > $T/r2g_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 $T/r2g_pytest.log
timeout 200 python tools/kernel_bench.py attn > $T/r2g_kernel_bench_attn_ilv.log 2>&1; grep -E "fa_fwd|SDPA" $T/r2g_kernel_bench_attn_ilv.log | cut -c1-200
AVJ_ATTN_ILV=0 timeout 200 python tools/kernel_bench.py attn > $T/r2g_kernel_bench_attn_noilv.log 2>&1; grep -E "fa_fwd" $T/r2g_kernel_bench_attn_noilv.log | cut -c1-200
if [ $rc -eq 0 ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2g_prof_dump.csv > $T/r2g_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2g_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity']['ok'], d['parity']['grad_rel'])"
fi
