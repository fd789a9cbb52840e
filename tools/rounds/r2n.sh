#!/bin/bash
# round 2, call n: LayerNorm backward with up-front loads -- numerics, microbench, bench
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "layernorm or row" > $T/r2n_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 $T/r2n_pytest.log | cut -c1-300
timeout 200 python tools/kernel_bench.py misc > $T/r2n_kernel_bench_misc.log 2>&1; grep -E "layernorm" $T/r2n_kernel_bench_misc.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --prof-dump $T/r2n_prof_dump.csv > $T/r2n_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2n_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()}, d['parity']['ok'], d['clocks'])"
