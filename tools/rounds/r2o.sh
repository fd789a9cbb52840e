#!/bin/bash
# round 2, call o: attentive pooler (few-query cross-attention kernel) vs oracle; LayerNorm-backward cost attribution
mkdir -p gpurun_out
T=gpurun_out
timeout 600 python -m pytest tests/test_pooler.py -m gpu -q --timeout 300 > $T/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 $T/r2o_pytest.log | cut -c1-300
timeout 200 python tools/micro/ln_bwd_variants.py > $T/r2o_ln_variants.log 2>&1; tail -6 $T/r2o_ln_variants.log | cut -c1-200
