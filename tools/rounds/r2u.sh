#!/bin/bash
# round 2, call u: im2col-free patch embedding (4-D TMA boxes out of the clip, tf32 tcgen05.mma): parity, step tests, bench A/B
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "patch" > $T/r2u_pytest_patch.log 2>&1
echo "pytest patch rc=$?"; tail -15 $T/r2u_pytest_patch.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > $T/r2u_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -8 $T/r2u_pytest_all.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --prof-dump $T/r2u_prof_dump.csv > $T/r2u_bench.json 2> $T/r2u_bench.err
echo "bench rc=$?"; tail -2 $T/r2u_bench.err | cut -c1-300; cut -c1-300 $T/r2u_bench.json
python tools/step_breakdown.py $T/r2u_prof_dump.csv > $T/r2u_step_breakdown.txt 2>&1; grep -i "patchify\|37632\|x1536\|x256 \|total" $T/r2u_step_breakdown.txt | head
AVJ_PATCH_EMBED_TMA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity > $T/r2u_bench_patchify.json 2> $T/r2u_bench_patchify.err
echo "bench patchify rc=$?"; cut -c1-300 $T/r2u_bench_patchify.json
