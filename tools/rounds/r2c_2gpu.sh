#!/bin/bash
# round 2, call c (2 GPUs): NCCL data-parallel parity test (TrainStep + GradSync vs single-GPU full batch), 2-GPU bench with
# comm_exposed_ms, overlap on/off, same-masks A/B
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q -s --timeout 900 > $T/r2c_pytest_nccl.log 2>&1
echo "nccl pytest rc=$?"; grep -E "grad_rel_err|passed|failed|ok" $T/r2c_pytest_nccl.log | cut -c1-300 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for tag in default sameMasks noOverlap; do
  extra=""; envs=""
  [ $tag = sameMasks ] && extra="--same-masks"
  if [ $tag = noOverlap ]; then export AVJ_DDP_OVERLAP=0; else unset AVJ_DDP_OVERLAP; fi
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 $extra > $T/r2c_bench2_$tag.log 2>&1
  echo "bench2 $tag rc=$?"; grep '"metric"' $T/r2c_bench2_$tag.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d.get('comm_exposed_ms'))"
done
grep -iE "NVLS|nvls" $T/r2c_bench2_default.log | head -5
