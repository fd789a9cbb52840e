#!/bin/bash
# round 2, call zk: fp32-output GEMM epilogue with the addend loads one chunk ahead: parity, per-shape times, step
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -m gpu -q --timeout 300 -k "gemm or patch or step" > $T/r2zk_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $T/r2zk_pytest.log | cut -c1-300
timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -v cublas | grep "proj\|fc2\|wgrad" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity --no-cpu-baseline --prof-dump $T/r2zk_prof.csv > $T/r2zk_bench.json 2> $T/r2zk_bench.err
echo "bench rc=$?"; grep "\[bench\]" $T/r2zk_bench.err
python tools/step_breakdown.py $T/r2zk_prof.csv 2>/dev/null | grep "f32out\|total\|gemm$" | head -24
