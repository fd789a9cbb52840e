#!/bin/bash
# round 25 (1 GPU): new GPU tests only (eval helpers, timing dump, staged loss read, alternative kernel paths)
mkdir -p gpurun_out
timeout 112 python -m pytest tests/test_eval_helpers_gpu.py tests/test_fallback_paths_gpu.py -q -m gpu --timeout 100 > gpurun_out/pytest25.log 2>&1
echo "rc=$?"; tail -40 gpurun_out/pytest25.log | cut -c1-600
