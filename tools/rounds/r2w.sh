#!/bin/bash
# round 2, call w: the lines kept in profiles/ for the final build -- default bench (with parity block, reference-on-GPU arm and
# cpu_baseline), reference arm, ViT-B / ViT-H / frozen-forward configs, smoke, step breakdown
mkdir -p gpurun_out
T=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $T/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $T/r2w_smoke.log | cut -c1-300
timeout 900 python bench.py --prof-dump $T/r2w_prof_dump.csv > $T/r2w_bench.json 2> $T/r2w_bench.err; echo "bench rc=$?"; tail -2 $T/r2w_bench.err | cut -c1-200
python tools/step_breakdown.py $T/r2w_prof_dump.csv > $T/r2w_step_breakdown.txt 2>&1; head -12 $T/r2w_step_breakdown.txt
timeout 900 python bench.py > $T/r2w_bench_noprof.json 2> $T/r2w_bench_noprof.err; echo "bench (no prof) rc=$?"; cut -c1-250 $T/r2w_bench_noprof.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $T/r2w_bench_reference.json 2> $T/r2w_bench_reference.err; echo "ref rc=$?"; cut -c1-400 $T/r2w_bench_reference.json
timeout 600 python bench.py --model vit_base --batch 32 --steps 10 --warmup 3 > $T/r2w_bench_vitb_b32.json 2> $T/r2w_bench_vitb.err; echo "vitb rc=$?"; cut -c1-250 $T/r2w_bench_vitb_b32.json
timeout 600 python bench.py --model vit_huge --batch 24 --steps 5 --warmup 3 --no-reference-gpu > $T/r2w_bench_vith_b24.json 2> $T/r2w_bench_vith.err; echo "vith rc=$?"; cut -c1-250 $T/r2w_bench_vith_b24.json
timeout 600 python bench.py --frozen-forward --batch 64 --steps 10 --warmup 3 > $T/r2w_bench_frozen_b64.json 2> $T/r2w_bench_frozen.err; echo "frozen rc=$?"; cut -c1-250 $T/r2w_bench_frozen_b64.json
