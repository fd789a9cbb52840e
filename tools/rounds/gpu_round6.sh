#!/bin/bash
# round 6: polynomial GELU, dual-path GEMM epilogue + L2 prefetch, library-side profiling
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest6.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest6.log
for mode in 0 1 2; do
  AVJ_EPI_MODE=$mode timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1e_mode$mode.log 2>&1
  echo "== epilogue mode $mode"; grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1e_mode$mode.log | cut -c1-190
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1e.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1e.log | cut -c1-2500
CASES="gemm_qkv gemm_proj gemm_fc1 gemm_dact gemm_pred_fc1 gemm_square"
timeout 300 python tools/ncu_cases.py $CASES > gpurun_out/ncu_cases_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -c 12 -o gpurun_out/prof_gemm_r1e -f \
  python tools/ncu_cases.py $CASES > gpurun_out/ncu_gemm_r1e.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm_r1e.log
