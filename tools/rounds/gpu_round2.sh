#!/bin/bash
# second GPU session: tests with the tensor-core attention, ViT-L bench, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest2.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest2.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_vitl.log 2>&1
echo "bench rc=$?"; tail -2 gpurun_out/bench_vitl.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 400 -c 3 -o gpurun_out/prof_gemm_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
