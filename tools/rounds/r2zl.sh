#!/bin/bash
# round 2, call zl: single-pass attention backward with dQ formed once per pair of steps (M = 128)
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "attention" > $T/r2zl_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $T/r2zl_pytest.log | cut -c1-400
timeout 300 python tools/kernel_bench.py attn 2>&1 | grep "fa_bwd" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity --no-cpu-baseline > $T/r2zl_bench.json 2> $T/r2zl_bench.err
echo "bench rc=$?"; grep "\[bench\]" $T/r2zl_bench.err
