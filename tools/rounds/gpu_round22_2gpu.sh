#!/bin/bash
# round 22 (2 GPUs): why is the e2e loop slower than the resident loop at N=2 (and not at N=1)?
mkdir -p gpurun_out
T=gpurun_out
run() { timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $T/bench_r1u_2gpu_$2.log 2>&1; }
AVJ_E2E_NOSYNC=1 run 29531 nosync 90; echo "nosync rc=$?"; tail -1 $T/bench_r1u_2gpu_nosync.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
run 29532 default 90; echo "default rc=$?"; tail -1 $T/bench_r1u_2gpu_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
