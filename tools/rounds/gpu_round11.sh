#!/bin/bash
# round 11: 16-warp direct epilogues, fused bias-grad colsums; A/B of 2-CTA x EW16 on the full step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest11.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest11.log
timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1j.log 2>&1
grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1j.log | cut -c1-190
for cfg in "2 1" "0 1" "2 0" "0 0" "2 1" "0 1"; do
  set -- $cfg
  AVJ_GEMM_2CTA=$1 AVJ_GEMM_EW16=$2 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1j_2cta$1_ew$2.log 2>&1
  echo "== 2CTA=$1 EW16=$2 rc=$?"; tail -1 gpurun_out/bench_r1j_2cta$1_ew$2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
done
