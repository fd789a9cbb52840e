#!/bin/bash
# round 2, call x: pooled activation arenas (cudaMalloc calls inside timed steps), A/B; ncu --set full of the predictor attention
# kernels at the exact sequence length of the profiled bench step
mkdir -p gpurun_out
T=gpurun_out
for pool in 1 0 1 0; do
AVJ_ARENA_POOL=$pool timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity --no-cpu-baseline > $T/r2x_bench_pool$pool.json 2> $T/r2x_bench_pool$pool.err
echo "pool=$pool rc=$?"; grep "\[bench\]" $T/r2x_bench_pool$pool.err
done
timeout 100 python tools/ncu_cases.py attn_pred_step > $T/r2x_plain.log 2>&1 &&
timeout 500 ncu --set full --import-source on --clock-control none -k regex:fa_ -o $T/r2x_prof_attn_pred_step -f python tools/ncu_cases.py attn_pred_step > $T/r2x_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2x_ncu.log
