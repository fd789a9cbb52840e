#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_eval_helpers_gpu.py -q -m gpu --timeout 50 -k "prof_dump or host_read" > gpurun_out/pytest26.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/pytest26.log | cut -c1-300
