#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/attn_check.py <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import torch, kernel_checks as kc
for case in [(1,2,128,1,64),(1,1,256,2,64),(1,2,257,4,64),(1,1,130,2,24),(1,2,1664,3,64),(1,3,1216,2,24),(1,1,40,2,32),(1,2,100,2,48),(1,1,96,2,16)]:
    try:
        ok, err = kc.check_attention(*case); torch.cuda.synchronize()
        print(json.dumps(dict(case=case, ok=bool(ok), err=err)), flush=True)
    except Exception as e:
        print(json.dumps(dict(case=case, ok=False, exc=repr(e)[:300])), flush=True)
PY
timeout 300 python /tmp/attn_check.py > gpurun_out/attn_umma_check.log 2>&1
echo "attn umma check rc=$?"; cat gpurun_out/attn_umma_check.log | grep -v Warn | tail -12
timeout 900 python tools/kernel_bench.py all > gpurun_out/kernel_bench_r1b.log 2>&1
echo "kernel bench rc=$?"; grep -E "^\{" gpurun_out/kernel_bench_r1b.log | cut -c1-260
AVJ_ATTN_FWD=mma timeout 300 python tools/kernel_bench.py attn > gpurun_out/kernel_bench_attn_mma.log 2>&1
grep -E "fa_fwd" gpurun_out/kernel_bench_attn_mma.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest3.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest3.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1b.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1b.log | cut -c1-900
