#!/bin/bash
# round 2, call zb: single-pass attention backward with the dQ read-out behind the operand hand-over
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "attention" > $T/r2zb_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $T/r2zb_pytest.log | cut -c1-300
timeout 300 python tools/kernel_bench.py attn 2>&1 | grep "fa_bwd\|fa_fwd" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity --no-cpu-baseline > $T/r2zb_bench.json 2> $T/r2zb_bench.err
echo "bench rc=$?"; grep "\[bench\]" $T/r2zb_bench.err
