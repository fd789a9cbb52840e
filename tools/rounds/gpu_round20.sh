#!/bin/bash
# round 20 (1 GPU): final validation -- full GPU test suite, smoke(), bench line with the per-shape roofline
mkdir -p gpurun_out
T=gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q --timeout 150 > $T/pytest20.log 2>&1
echo "pytest all rc=$?"; tail -3 $T/pytest20.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $T/smoke20.log 2>&1
echo "smoke rc=$?"; tail -2 $T/smoke20.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1s.csv > $T/bench_r1s.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1s.log | cut -c1-3000
