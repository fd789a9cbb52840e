#!/bin/bash
# round 2, call b: merged two-mask variable-length schedule (one stack pass for both masks) -- full GPU suite, bench A/B
mkdir -p gpurun_out
T=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $T/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 $T/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2b_prof_dump.csv > $T/r2b_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2b_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'launches/step', d['gpu_launches_per_step'], {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity'], d['clocks'])"
python tools/step_breakdown.py $T/r2b_prof_dump.csv 70 > $T/r2b_step_breakdown.txt 2>&1
AVJ_MERGE_MASKS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-parity > $T/r2b_bench_nomerge.log 2>&1
echo "bench(no merge) rc=$?"; tail -1 $T/r2b_bench_nomerge.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
