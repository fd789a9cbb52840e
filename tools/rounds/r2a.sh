#!/bin/bash
# round 2, call a: validate the round-2 groundwork on one B200 -- full GPU suite (BASELINE-config parity, branches, logging
# statistics), smoke(), the bench line with parity + reference-on-GPU + live-reference CPU legs, library columns of the
# per-kernel microbenchmarks.
mkdir -p gpurun_out
T=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $T/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > $T/r2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $T/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $T/r2a_smoke.log 2>&1
echo "smoke rc=$?"; tail -4 $T/r2a_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --prof-dump $T/r2a_prof_dump.csv > $T/r2a_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2a_bench.log | cut -c1-3000
python tools/step_breakdown.py $T/r2a_prof_dump.csv 40 > $T/r2a_step_breakdown.txt 2>&1
timeout 200 python tools/kernel_bench.py gemm > $T/r2a_kernel_bench_gemm.log 2>&1
timeout 200 python tools/kernel_bench.py attn > $T/r2a_kernel_bench_attn.log 2>&1
echo "kernel_bench done"; tail -3 $T/r2a_kernel_bench_attn.log | cut -c1-300
