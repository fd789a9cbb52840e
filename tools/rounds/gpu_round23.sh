#!/bin/bash
# round 23 (1 GPU): loss scalars staged to the host behind the forward -- step tests + bench
mkdir -p gpurun_out
T=gpurun_out
timeout 200 python -m pytest tests/test_step_gpu.py -m gpu -x -q --timeout 150 > $T/pytest23.log 2>&1
echo "pytest step rc=$?"; tail -2 $T/pytest23.log
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $T/bench_r1v.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['loss'], 'e2e', d['e2e'])"
