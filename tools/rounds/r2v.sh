#!/bin/bash
# round 2, call v: target encoder on a second stream (A/B), polynomial exp2 share for long sequences
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > $T/r2v_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -4 $T/r2v_pytest_all.log | cut -c1-300
for rep in 1 2; do
for ts in 1 0; do
AVJ_TARGET_STREAM=$ts timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity > $T/r2v_bench_ts${ts}_$rep.json 2> $T/r2v_bench_ts${ts}_$rep.err
echo "target stream=$ts rc=$?"; python -c "
import json; d=json.loads(open('$T/r2v_bench_ts${ts}_$rep.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('frac_of_bf16_peak'))"
done; done
