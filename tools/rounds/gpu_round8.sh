#!/bin/bash
# round 8: batched residual loads, tanh GELU, LN bwd rewrite, colsum rewrite; A/B of the L2 prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest8.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest8.log
for pf in 1 0; do
  AVJ_GEMM_L2PF=$pf timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1g_pf$pf.log 2>&1
  echo "== L2 prefetch $pf"; grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1g_pf$pf.log | cut -c1-190
done
timeout 300 python tools/kernel_bench.py misc > gpurun_out/kernel_bench_misc_r1g.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_misc_r1g.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1g.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1g.log | cut -c1-3000
