#!/bin/bash
# round 12: attention backward row statistics precomputed (no loader-side global loads), TMA loads for hd <= 32
# (SWIZZLE_64B boxes + pad zeroing of the stationary tiles), per-launch timing dump of one full step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest12.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest12.log
if [ $rc -ne 0 ]; then
  AVJ_GEMM_TMA_STORE=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 -k gemm > gpurun_out/pytest12_nots.log 2>&1
  echo "pytest gemm (no TMA store) rc=$?"; tail -5 gpurun_out/pytest12_nots.log
  AVJ_ATTN_TMA=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 300 -k attention > gpurun_out/pytest12_notma.log 2>&1
  echo "pytest attention (no TMA) rc=$?"; tail -5 gpurun_out/pytest12_notma.log
fi
for tma in 1 0; do
  AVJ_ATTN_TMA=$tma timeout 300 python tools/kernel_bench.py attn > gpurun_out/kernel_bench_attn_r1k_tma$tma.log 2>&1
  echo "== attn AVJ_ATTN_TMA=$tma"; grep -E "fa_" gpurun_out/kernel_bench_attn_r1k_tma$tma.log | cut -c1-200
done
for ts in 1 0; do
  AVJ_GEMM_TMA_STORE=$ts timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1k_ts$ts.log 2>&1
  echo "== gemm AVJ_GEMM_TMA_STORE=$ts"; grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1k_ts$ts.log | cut -c1-190
done
timeout 300 python tools/kernel_bench.py misc > gpurun_out/kernel_bench_misc_r1k.log 2>&1
cat gpurun_out/kernel_bench_misc_r1k.log | cut -c1-200
for cfg in "2 1 1" "2 1 0" "0 1 1" "2 0 1"; do
  set -- $cfg
  AVJ_GEMM_2CTA=$1 AVJ_GEMM_EW16=$2 AVJ_GEMM_TMA_STORE=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump gpurun_out/prof_dump_r1k_2cta$1_ew$2_ts$3.csv > gpurun_out/bench_r1k_2cta$1_ew$2_ts$3.log 2>&1
  echo "== 2CTA=$1 EW16=$2 TS=$3 rc=$?"; tail -1 gpurun_out/bench_r1k_2cta$1_ew$2_ts$3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
done
python tools/step_breakdown.py gpurun_out/prof_dump_r1k_2cta2_ew1_ts1.csv 45
