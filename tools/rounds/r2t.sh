#!/bin/bash
# round 2, call t (2 GPUs): dynamic vs static GEMM unit schedule under the overlapped gradient all-reduce (A/B on one box)
mkdir -p gpurun_out
T=gpurun_out
run() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-reference-gpu > $T/r2t_${tag}.json 2> $T/r2t_${tag}.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('$T/r2t_${tag}.json').read().strip().splitlines()[-1])
    print('   ', d.get('ms_per_step'), d.get('comm_exposed_ms', {}).get('resident'), d.get('e2e', {}).get('ms_per_step'))
except Exception as e:
    print('no line', e)
PY
}
run dyn A=1
run static AVJ_GEMM_DYNAMIC=0
run dyn_nooverlap AVJ_DDP_OVERLAP=0
run static_nooverlap AVJ_GEMM_DYNAMIC=0 AVJ_DDP_OVERLAP=0
run dyn2 A=1
run static2 AVJ_GEMM_DYNAMIC=0
