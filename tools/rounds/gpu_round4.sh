#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest4.log 2>&1
echo "pytest kernels rc=$?"; tail -3 gpurun_out/pytest4.log
timeout 300 python /dev/stdin > gpurun_out/attn_umma_check2.log 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import torch, kernel_checks as kc
for case in [(1,2,128,1,64),(1,2,257,4,64),(1,1,130,2,24),(1,2,1664,3,64),(1,3,1216,2,24),(1,2,100,2,48),(1,24,384,16,64)]:
    try:
        ok, err = kc.check_attention(*case); torch.cuda.synchronize()
        print(json.dumps(dict(case=case, ok=bool(ok), err=err)), flush=True)
    except Exception as e:
        print(json.dumps(dict(case=case, ok=False, exc=repr(e)[:300])), flush=True)
PY
grep -E "^\{" gpurun_out/attn_umma_check2.log
timeout 600 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1c.log 2>&1
grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1c.log | cut -c1-220
timeout 300 python tools/kernel_bench.py attn > gpurun_out/kernel_bench_attn_r1c.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_attn_r1c.log | cut -c1-200
AVJ_ATTN_TMA=0 timeout 300 python tools/kernel_bench.py attn 2>&1 | grep fa_fwd | cut -c1-200
