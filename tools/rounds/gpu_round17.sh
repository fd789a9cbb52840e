#!/bin/bash
# round 17: P / dS through tensor memory (TS-form tcgen05.mma), programmatic dependent launch; fallbacks by env
mkdir -p gpurun_out
T=gpurun_out
try() { timeout $1 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k "$2" > $T/pytest17_$3.log 2>&1; }
try 200 attention attn_default; rc=$?; echo "attention default rc=$rc"; tail -3 $T/pytest17_attn_default.log
if [ $rc -ne 0 ]; then
  export AVJ_ATTN_TMEM_P=0; try 150 attention attn_nots; rc=$?; echo "attention TMEM_P=0 rc=$rc"; tail -3 $T/pytest17_attn_nots.log
  if [ $rc -ne 0 ]; then
    export AVJ_PDL=0; try 150 attention attn_nots_nopdl; rc=$?; echo "attention TMEM_P=0 PDL=0 rc=$rc"
    if [ $rc -ne 0 ]; then echo "attention broken"; exit 1; fi
    unset AVJ_ATTN_TMEM_P; try 150 attention attn_nopdl; rc=$?; echo "attention PDL=0 (TMEM_P on) rc=$rc"
    if [ $rc -ne 0 ]; then export AVJ_ATTN_TMEM_P=0; fi
  fi
fi
try 150 "gemm or layernorm" gemm_ln; rc=$?; echo "gemm+ln rc=$rc"; tail -2 $T/pytest17_gemm_ln.log
if [ $rc -ne 0 ] && [ -z "$AVJ_PDL" ]; then export AVJ_PDL=0; try 150 "gemm or layernorm" gemm_ln_nopdl; echo "gemm+ln PDL=0 rc=$?"; fi
echo "ENV: TMEM_P=$AVJ_ATTN_TMEM_P PDL=$AVJ_PDL"
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > $T/pytest17.log 2>&1
echo "pytest all rc=$?"; tail -3 $T/pytest17.log
timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1p.log 2>&1
echo "== attn"; grep -E "fa_" $T/kernel_bench_attn_r1p.log | cut -c1-200
if [ -z "$AVJ_ATTN_TMEM_P" ]; then
  AVJ_ATTN_TMEM_P=0 timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1p_nots.log 2>&1
  echo "== attn TMEM_P=0"; grep -E "fa_" $T/kernel_bench_attn_r1p_nots.log | cut -c1-200
fi
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1p.csv > $T/bench_r1p.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
if [ -z "$AVJ_PDL" ]; then
  AVJ_PDL=0 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $T/bench_r1p_nopdl.log 2>&1
  echo "== bench PDL=0 rc=$?"; tail -1 $T/bench_r1p_nopdl.log | cut -c1-330
fi
