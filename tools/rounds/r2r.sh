#!/bin/bash
# round 2, call r (2 GPUs): does the all-reduce's SM occupancy stretch the persistent GEMMs?  NCCL CTA cap x GEMM grid cap
mkdir -p gpurun_out
T=gpurun_out
run() {  # tag, env...
  local tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-reference-gpu > $T/r2r_${tag}.json 2> $T/r2r_${tag}.err
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('$T/r2r_${tag}.json').read().strip().splitlines()[-1])
    print('   ', d.get('ms_per_step'), d.get('comm_exposed_ms', {}).get('resident'), d.get('e2e', {}).get('ms_per_step'))
except Exception as e:
    print('no line', e)
PY
}
run base A=1
run nooverlap AVJ_DDP_OVERLAP=0
run cta4 NCCL_MAX_CTAS=4
run cta4_sm140 NCCL_MAX_CTAS=4 AVJ_GEMM_SMS=140
run cta8_sm132 NCCL_MAX_CTAS=8 AVJ_GEMM_SMS=132
run cta2_sm144 NCCL_MAX_CTAS=2 AVJ_GEMM_SMS=144
run sm140 AVJ_GEMM_SMS=140
