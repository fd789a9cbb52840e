#!/bin/bash
# round 2, call y (8 GPUs): BASELINE config 4 (ViT-H/16, predictor depth 12) on 8 GPUs, and the 1-GPU ViT-L line of the same box
mkdir -p gpurun_out
T=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
  bench.py --gpus 8 --model vit_huge --batch 24 --steps 8 --warmup 3 --no-parity --no-reference-gpu > $T/r2y_bench_vith_n8.json 2> $T/r2y_bench_vith_n8.err
echo "vith n8 rc=$?"; grep "\[bench\]" $T/r2y_bench_vith_n8.err; cut -c1-300 $T/r2y_bench_vith_n8.json
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity --no-reference-gpu --no-cpu-baseline > $T/r2y_bench_vitl_n1.json 2> $T/r2y_bench_vitl_n1.err
echo "vitl n1 rc=$?"; grep "\[bench\]" $T/r2y_bench_vitl_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 \
  bench.py --gpus 8 --steps 10 --warmup 3 --no-parity --no-reference-gpu > $T/r2y_bench_vitl_n8.json 2> $T/r2y_bench_vitl_n8.err
echo "vitl n8 rc=$?"; grep "\[bench\]" $T/r2y_bench_vitl_n8.err; cut -c1-200 $T/r2y_bench_vitl_n8.json
