#!/bin/bash
# round 2, call p: tcgen05 attention for head_dim 80..128 (two 64-column halves): parity, kernel times, ViT-H step
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "attention" > $T/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $T/r2p_pytest.log | cut -c1-300
timeout 300 python tools/kernel_bench.py attn_h > $T/r2p_attn_h.log 2>&1; cat $T/r2p_attn_h.log | cut -c1-250
AVJ_ATTN_FWD=mma AVJ_ATTN_BWD=mma timeout 300 python tools/kernel_bench.py attn_h > $T/r2p_attn_h_mma.log 2>&1; cat $T/r2p_attn_h_mma.log | cut -c1-250
timeout 600 python bench.py --model vit_huge --batch 24 --steps 5 --warmup 3 --no-reference-gpu > $T/r2p_bench_vith.json 2> $T/r2p_bench_vith.err
echo "bench rc=$?"; tail -3 $T/r2p_bench_vith.err | cut -c1-300; cut -c1-1500 $T/r2p_bench_vith.json
