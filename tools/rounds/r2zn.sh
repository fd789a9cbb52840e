#!/bin/bash
# round 2, call zn: patch-embedding weight gradient with the operand gathered by the GEMM producer (no patch matrix in backward either)
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_step_configs_gpu.py -m gpu -q --timeout 500 -k "patch or step_matches or bf16 or fp32" -x > $T/r2zn_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $T/r2zn_pytest.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-cpu-baseline --prof-dump $T/r2zn_prof.csv > $T/r2zn_bench.json 2> $T/r2zn_bench.err
echo "bench rc=$?"; grep "\[bench\]" $T/r2zn_bench.err
python tools/step_breakdown.py $T/r2zn_prof.csv 2>/dev/null | grep -i "patchify\|x1536\|x256 \|1024x1536\|total" | head
python -c "
import json; d=json.loads(open('$T/r2zn_bench.json').read().strip().splitlines()[-1]); print(d['parity']['ok'], d['parity']['grad_rel'], d['parity']['worst_family'], d['parity']['worst_family_rel'])"
