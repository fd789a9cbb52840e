#!/bin/bash
# round 24 (2 GPUs): e2e loop with the loss scalars staged behind the forward
mkdir -p gpurun_out
T=gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $T/bench_r1w_2gpu.log 2>&1
echo "rc=$?"; tail -1 $T/bench_r1w_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['loss'], 'e2e', d['e2e'], d['clocks'])"
