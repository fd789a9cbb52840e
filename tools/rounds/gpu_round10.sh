#!/bin/bash
# round 10: CTA-pair GEMM (tcgen05 cta_group::2), DevicePrefetcher in the e2e loop
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 -k gemm > gpurun_out/pytest10_gemm.log 2>&1
rc=$?; echo "pytest gemm (2-CTA auto) rc=$rc"; tail -6 gpurun_out/pytest10_gemm.log
if [ $rc -ne 0 ]; then
  AVJ_GEMM_2CTA=0 timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 120 -k gemm > gpurun_out/pytest10_gemm_1cta.log 2>&1
  echo "pytest gemm (1-CTA) rc=$?"; tail -3 gpurun_out/pytest10_gemm_1cta.log
  exit 0
fi
for m in 2 0; do
  AVJ_GEMM_2CTA=$m timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1i_2cta$m.log 2>&1
  echo "== AVJ_GEMM_2CTA=$m"; grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1i_2cta$m.log | cut -c1-190
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest10.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest10.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1i.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1i.log | cut -c1-3000
AVJ_GEMM_2CTA=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1i_1cta.log 2>&1
echo "bench 1cta rc=$?"; tail -1 gpurun_out/bench_vitl_r1i_1cta.log | cut -c1-400
