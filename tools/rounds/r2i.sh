#!/bin/bash
# round 2, call i: single-pass attention backward for head_dim <= 32 (dQ via M=64 MMA + TMA reduce-add) -- numerics, microbench, bench
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "attention" > $T/r2i_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -8 $T/r2i_pytest.log | cut -c1-400
timeout 200 python tools/kernel_bench.py attn > $T/r2i_kernel_bench_attn_sp.log 2>&1; grep -E "fa_bwd" $T/r2i_kernel_bench_attn_sp.log | cut -c1-200
AVJ_ATTN_BWD_SP=0 timeout 200 python tools/kernel_bench.py attn > $T/r2i_kernel_bench_attn_nosp.log 2>&1; grep -E "fa_bwd" $T/r2i_kernel_bench_attn_nosp.log | cut -c1-200
if [ $rc -eq 0 ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2i_prof_dump.csv > $T/r2i_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2i_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity']['ok'], d['parity']['grad_rel'])"
fi
