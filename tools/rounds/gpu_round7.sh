#!/bin/bash
# round 7: setmaxnreg warp-group register split (GEMM 384 thr, attention 256 thr x 2 CTAs/SM), specialised epilogues
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest7.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest7.log
timeout 300 python tools/kernel_bench.py gemm > gpurun_out/kernel_bench_gemm_r1f.log 2>&1
grep -E "gemm_umma" gpurun_out/kernel_bench_gemm_r1f.log | cut -c1-190
timeout 300 python tools/kernel_bench.py attn > gpurun_out/kernel_bench_attn_r1f.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_attn_r1f.log | cut -c1-200
timeout 300 python tools/kernel_bench.py misc > gpurun_out/kernel_bench_misc_r1f.log 2>&1
grep -E "^\{" gpurun_out/kernel_bench_misc_r1f.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl_r1f.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_vitl_r1f.log | cut -c1-3000
CASES="attn_target attn_pred gemm_qkv gemm_proj gemm_fc1"
timeout 300 python tools/ncu_cases.py $CASES > gpurun_out/ncu_cases_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fa_|gemm_umma" -c 22 -o gpurun_out/prof_r1f -f \
  python tools/ncu_cases.py $CASES > gpurun_out/ncu_r1f.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r1f.log
