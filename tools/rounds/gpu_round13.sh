#!/bin/bash
# round 13: setmaxnreg pool fix for the 16-warp epilogue, TMA-store epilogue, attention row statistics + TMA for
# hd <= 32, unmasked softmax tiles.  Every stage has a tight timeout and falls back to the previous path by
# environment switch, so one hanging kernel cannot eat the whole GPU budget.
mkdir -p gpurun_out
T=gpurun_out
try_gemm() { timeout $1 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k gemm > $T/pytest13_gemm_$2.log 2>&1; }
try_gemm 240 default; rc=$?; echo "gemm default rc=$rc"; tail -2 $T/pytest13_gemm_default.log
if [ $rc -ne 0 ]; then
  export AVJ_GEMM_TMA_STORE=0; try_gemm 150 nots; rc=$?; echo "gemm TMA_STORE=0 rc=$rc"; tail -2 $T/pytest13_gemm_nots.log
  if [ $rc -ne 0 ]; then
    export AVJ_GEMM_EW16=0; try_gemm 150 nots_noew16; rc=$?; echo "gemm TMA_STORE=0 EW16=0 rc=$rc"; tail -2 $T/pytest13_gemm_nots_noew16.log
    if [ $rc -ne 0 ]; then echo "GEMM broken even with fallbacks"; exit 1; fi
    unset AVJ_GEMM_TMA_STORE; try_gemm 150 noew16; rc=$?; echo "gemm EW16=0 (TMA store on) rc=$rc"
    if [ $rc -ne 0 ]; then export AVJ_GEMM_TMA_STORE=0; fi
  fi
fi
try_attn() { timeout $1 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k attention > $T/pytest13_attn_$2.log 2>&1; }
try_attn 150 default; rc=$?; echo "attention default rc=$rc"; tail -2 $T/pytest13_attn_default.log
if [ $rc -ne 0 ]; then
  export AVJ_ATTN_TMA32=0; try_attn 150 notma32; rc=$?; echo "attention TMA32=0 rc=$rc"; tail -2 $T/pytest13_attn_notma32.log
  if [ $rc -ne 0 ]; then echo "attention broken"; exit 1; fi
fi
echo "ENV: TMA_STORE=$AVJ_GEMM_TMA_STORE EW16=$AVJ_GEMM_EW16 ATTN_TMA32=$AVJ_ATTN_TMA32"
timeout 500 python -m pytest tests -m gpu -x -q --timeout 200 > $T/pytest13.log 2>&1
echo "pytest all rc=$?"; tail -3 $T/pytest13.log
timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1l.log 2>&1
echo "== attn"; grep -E "fa_" $T/kernel_bench_attn_r1l.log | cut -c1-200
if [ -z "$AVJ_ATTN_TMA32" ]; then
  AVJ_ATTN_TMA32=0 timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1l_notma32.log 2>&1
  echo "== attn TMA32=0"; grep -E "predictor" $T/kernel_bench_attn_r1l_notma32.log | cut -c1-200
fi
timeout 150 python tools/kernel_bench.py gemm > $T/kernel_bench_gemm_r1l.log 2>&1
echo "== gemm"; grep -E "gemm_umma" $T/kernel_bench_gemm_r1l.log | cut -c1-190
if [ -z "$AVJ_GEMM_TMA_STORE" ]; then
  AVJ_GEMM_TMA_STORE=0 timeout 150 python tools/kernel_bench.py gemm > $T/kernel_bench_gemm_r1l_nots.log 2>&1
  echo "== gemm TMA_STORE=0"; grep -E "gemm_umma" $T/kernel_bench_gemm_r1l_nots.log | cut -c1-190
fi
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1l.csv > $T/bench_r1l.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1l.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
python tools/step_breakdown.py $T/prof_dump_r1l.csv 45
AVJ_GEMM_2CTA=0 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $T/bench_r1l_1cta.log 2>&1
echo "== bench 1cta rc=$?"; tail -1 $T/bench_r1l_1cta.log | cut -c1-330
