#!/bin/bash
# round 21 (2 GPUs): data-parallel bench over NCCL, plain all-reduce (default) then the overlapped one
mkdir -p gpurun_out
T=gpurun_out
run() { timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $T/bench_r1t_2gpu_$2.log 2>&1; }
run 29521 plain 140; echo "plain rc=$?"; tail -1 $T/bench_r1t_2gpu_plain.log | cut -c1-700
AVJ_DDP_OVERLAP=1 run 29522 overlap 125; echo "overlap rc=$?"; tail -1 $T/bench_r1t_2gpu_overlap.log | cut -c1-700
