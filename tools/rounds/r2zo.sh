#!/bin/bash
# round 2, call zo: wgrad gather with four chunks in flight
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -k "patch" -x > $T/r2zo_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 $T/r2zo_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-cpu-baseline --no-parity --prof-dump $T/r2zo_prof.csv > $T/r2zo_bench.json 2> $T/r2zo_bench.err
echo "bench rc=$?"; grep "\[bench\]" $T/r2zo_bench.err
python tools/step_breakdown.py $T/r2zo_prof.csv 2>/dev/null | grep -i "patchify\|1024x1536\|1024x256\|total" | head
