#!/bin/bash
# round 2, call e: LayerNorm-backward final reduction / colsum tuning / split-K wave model -- kernel tests, microbenchmarks,
# bench; BASELINE configs 2, 4, 5 bench lines; ncu --set full of the HBM-bound kernels (north-star: achieved GB/s + dram bytes)
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_branches_gpu.py -m gpu -q --timeout 300 -x > $T/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $T/r2e_pytest.log
timeout 200 python tools/kernel_bench.py misc > $T/r2e_kernel_bench_misc.log 2>&1; cat $T/r2e_kernel_bench_misc.log | cut -c1-200
timeout 200 python tools/kernel_bench.py gemm > $T/r2e_kernel_bench_gemm.log 2>&1; grep -E "wgrad" $T/r2e_kernel_bench_gemm.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-parity --prof-dump $T/r2e_prof_dump.csv > $T/r2e_bench.log 2>&1
echo "bench rc=$?"; tail -1 $T/r2e_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(v['ms'],v['achieved']) for k,v in d['roofline']['families'].items()})"
python tools/step_breakdown.py $T/r2e_prof_dump.csv 70 > $T/r2e_step_breakdown.txt 2>&1
timeout 900 python bench.py --model vit_base --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > $T/r2e_bench_vitb_b32.log 2>&1
echo "vit_base rc=$?"; tail -1 $T/r2e_bench_vitb_b32.log | cut -c1-400
timeout 900 python bench.py --model vit_huge --batch 24 --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > $T/r2e_bench_vith_b24.log 2>&1
echo "vit_huge rc=$?"; tail -1 $T/r2e_bench_vith_b24.log | cut -c1-400
timeout 600 python bench.py --frozen-forward --batch 64 --steps 10 --warmup 3 > $T/r2e_bench_frozen_b64.log 2>&1
echo "frozen rc=$?"; tail -1 $T/r2e_bench_frozen_b64.log | cut -c1-600
timeout 200 python tools/ncu_cases.py ln ln_ctx ln_pred colsum adamw gather loss > $T/r2e_ncu_plain.log 2>&1 &&
timeout 800 ncu --set full --import-source on --clock-control none -k regex:"layernorm|colsum|adamw|gather_rows|loss_kernel" -o $T/r2e_prof_hbm -f python tools/ncu_cases.py ln ln_ctx ln_pred colsum adamw gather loss > $T/r2e_ncu_hbm.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2e_ncu_hbm.log
