#!/bin/bash
# round 19 (2 GPUs): data-parallel bench over NCCL with the overlapped gradient all-reduce on and off
mkdir -p gpurun_out
T=gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $T/bench_r1r_2gpu_$2.log 2>&1; }
AVJ_DDP_OVERLAP=1 run 29511 overlap; echo "overlap rc=$?"; tail -1 $T/bench_r1r_2gpu_overlap.log | cut -c1-420
AVJ_DDP_OVERLAP=0 run 29512 plain; echo "plain rc=$?"; tail -1 $T/bench_r1r_2gpu_plain.log | cut -c1-420
grep -h -E "Error|error|Traceback|assert" $T/bench_r1r_2gpu_*.log | head -10
