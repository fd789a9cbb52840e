#!/bin/bash
# first GPU session: bring-up table, SIMT-only pipeline check, full pytest, smoke, small bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python tools/gpu_bringup.py > gpurun_out/bringup.log 2>&1
echo "bringup rc=$?"
AVJ_FORCE_SIMT=1 timeout 1200 python -m pytest tests/test_step_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_simt.log 2>&1
echo "pytest simt rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"
timeout 600 python bench.py --model vit_tiny --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny.log 2>&1
echo "bench tiny rc=$?"
tail -60 gpurun_out/bringup.log
tail -30 gpurun_out/pytest_simt.log
tail -30 gpurun_out/pytest.log
tail -5 gpurun_out/smoke.log
tail -3 gpurun_out/bench_tiny.log
