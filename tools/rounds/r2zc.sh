#!/bin/bash
# round 2, call zc: device-side mask sampling: parity tests, e2e A/B (host masks vs device masks), fallback-path tests
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_mask_collate_gpu.py tests/test_fallback_paths_gpu.py -m gpu -q --timeout 300 > $T/r2zc_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $T/r2zc_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity --no-cpu-baseline --device-masks > $T/r2zc_bench_devmasks.json 2> $T/r2zc_bench_devmasks.err
echo "bench device masks rc=$?"; grep "\[bench\]" $T/r2zc_bench_devmasks.err; tail -2 $T/r2zc_bench_devmasks.err | cut -c1-300
python -c "
import json; d=json.loads(open('$T/r2zc_bench_devmasks.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['config'].get('e2e_masks'))"
