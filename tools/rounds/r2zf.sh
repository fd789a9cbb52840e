#!/bin/bash
# round 2, call zf: patch-embedding kernel alone (dense and kept-token forms) and its ncu --set full capture
mkdir -p gpurun_out
T=gpurun_out
timeout 200 python tools/kernel_bench.py patch > $T/r2zf_patch.log 2>&1; cat $T/r2zf_patch.log | cut -c1-250
timeout 100 python tools/ncu_cases.py patch_embed > $T/r2zf_plain.log 2>&1 &&
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_umma_kernel -o $T/r2zf_prof_patch_embed -f python tools/ncu_cases.py patch_embed > $T/r2zf_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $T/r2zf_ncu.log
