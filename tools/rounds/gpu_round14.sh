#!/bin/bash
# round 14: TMEM-read / MUFU microbenchmark, FMA-pipe exp2 offload (AVJ_ATTN_POLY=1), ncu source-level capture of
# the predictor-shaped attention kernels, bench
mkdir -p gpurun_out
T=gpurun_out
timeout 60 tools/micro/tmem_bw > $T/tmem_bw.log 2>&1; echo "tmem_bw rc=$?"; cat $T/tmem_bw.log
AVJ_ATTN_POLY=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 100 -k attention > $T/pytest14_attn_poly.log 2>&1
echo "attention POLY rc=$?"; tail -3 $T/pytest14_attn_poly.log
for poly in 0 1; do
  AVJ_ATTN_POLY=$poly timeout 150 python tools/kernel_bench.py attn > $T/kernel_bench_attn_r1m_poly$poly.log 2>&1
  echo "== attn POLY=$poly"; grep -E "fa_" $T/kernel_bench_attn_r1m_poly$poly.log | cut -c1-200
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:fa_ -o $T/prof_attn_pred_r1m -f python tools/ncu_cases.py attn_pred > $T/ncu_attn_r1m.log 2>&1
echo "ncu rc=$?"; tail -3 $T/ncu_attn_r1m.log
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-dump $T/prof_dump_r1m.csv > $T/bench_r1m.log 2>&1
echo "== bench rc=$?"; tail -1 $T/bench_r1m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:v['ms'] for k,v in d['roofline']['families'].items()}, d['clocks'])"
AVJ_ATTN_POLY=1 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $T/bench_r1m_poly.log 2>&1
echo "== bench POLY rc=$?"; tail -1 $T/bench_r1m_poly.log | cut -c1-330
