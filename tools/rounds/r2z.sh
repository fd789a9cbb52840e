#!/bin/bash
# round 2, call z (2 GPUs): optimizer pipelined under the trailing gradient all-reduces: NCCL parity + A/B
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q --timeout 900 -s > $T/r2z_pytest_nccl.log 2>&1
echo "pytest nccl rc=$?"; grep -a "world\|passed\|failed\|Error" $T/r2z_pytest_nccl.log | cut -c1-250 | tail -12
run() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-reference-gpu > $T/r2z_${tag}.json 2> $T/r2z_${tag}.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('$T/r2z_${tag}.json').read().strip().splitlines()[-1])
    print('   ', d.get('ms_per_step'), d.get('comm_exposed_ms', {}).get('resident'), d.get('e2e', {}).get('ms_per_step'), d.get('comm_exposed_ms', {}).get('e2e'))
except Exception as e:
    print('no line', e)
PY
}
run pipe A=1
run plain AVJ_DDP_PIPELINE_OPT=0
run pipe2 A=1
run plain2 AVJ_DDP_PIPELINE_OPT=0
