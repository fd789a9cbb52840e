#!/bin/bash
# round 2, call zg: default bench line of the final build (with the full-size first-step loss probe against the reference on this GPU)
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python bench.py > $T/r2zg_bench.json 2> $T/r2zg_bench.err; echo "bench rc=$?"; grep "\[bench\]" $T/r2zg_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zg_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['frac_of_bf16_peak'])
print(d['reference_gpu'].get('full_size_parity'))
print({k: d['reference_gpu'][k] for k in ('with_loggers', 'no_loggers', 'speedup_vs_no_loggers', 'speedup_vs_with_loggers')})
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
print(d['parity']['ok'], d['parity']['grad_rel'], d['cpu_baseline']['value'])
PY
