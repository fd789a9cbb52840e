#!/bin/bash
# round 9: N-stage attention rings, 64-byte-pitch tiles (SWIZZLE_64B) for hd <= 32, faster delta kernel
mkdir -p gpurun_out
timeout 300 python /dev/stdin > gpurun_out/attn_check9.log 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import torch, kernel_checks as kc
for case in [(1,1,64,1,64),(1,2,257,4,64),(1,1,130,2,24),(1,2,375,3,64),(1,3,1216,2,24),(1,2,1664,3,64),(1,1,300,2,32),(1,1,96,2,16),(1,1,129,1,8),(1,2,100,2,48)]:
    try:
        ok, err = kc.check_attention(*case); torch.cuda.synchronize()
        print(json.dumps(dict(case=case, ok=bool(ok), err=err, parts=kc.LAST_ATTENTION_ERRORS)), flush=True)
    except Exception as e:
        print(json.dumps(dict(case=case, ok=False, exc=repr(e)[:300])), flush=True)
PY
echo "attn check rc=$?"; grep -E "^\{" gpurun_out/attn_check9.log | cut -c1-330
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest9.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest9.log
timeout 300 python tools/kernel_bench.py attn > gpurun_out/kernel_bench_attn_r1h.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_attn_r1h.log | cut -c1-200
timeout 300 python tools/kernel_bench.py misc > gpurun_out/kernel_bench_misc_r1h.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_misc_r1h.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1h.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1h.log | cut -c1-3000
