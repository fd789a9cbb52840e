#!/bin/bash
# round 16: two math warpgroups in the attention backward (A/B against one), ncu --set full of three GEMM cases
mkdir -p gpurun_out
T=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k attention > $T/pytest16_attn.log 2>&1
rc=$?; echo "attention MW=2 rc=$rc"; tail -3 $T/pytest16_attn.log
if [ $rc -ne 0 ]; then export AVJ_ATTN_BWD_MW=1; echo "falling back to MW=1"; fi
for mw in 2 1; do
  AVJ_ATTN_BWD_MW=$mw timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1o_mw$mw.log 2>&1
  echo "== attn MW=$mw"; grep -E "fa_bwd" $T/kernel_bench_attn_r1o_mw$mw.log | cut -c1-200
done
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1o.csv > $T/bench_r1o.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1o.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
timeout 100 python tools/ncu_cases.py gemm_fc1 gemm_pred_qkv gemm_pred_dact gemm_dact > $T/ncu_cases_plain_r1o.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:gemm_umma -o $T/prof_gemm_r1o -f python tools/ncu_cases.py gemm_fc1 gemm_pred_qkv gemm_pred_dact gemm_dact > $T/ncu_gemm_r1o.log 2>&1
echo "ncu rc=$?"; tail -3 $T/ncu_gemm_r1o.log
