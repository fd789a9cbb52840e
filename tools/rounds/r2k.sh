#!/bin/bash
# round 2, call k: where does the single-pass backward lose time?  (timing-only debug switches)
mkdir -p gpurun_out
T=gpurun_out
for d in 0 1 2; do
AVJ_ATTN_BWD_DBG=$d timeout 200 python tools/kernel_bench.py attn > $T/r2k_attn_dbg$d.log 2>&1; echo "dbg=$d"; grep -E "fa_bwd" $T/r2k_attn_dbg$d.log | grep predictor | cut -c1-200
done
