#!/bin/bash
# round 2, call q (8 GPUs): weak-scaling lines N = 8, 4, 2 with comm_exposed_ms; per-rank vs same-seed masks A/B at N = 8
mkdir -p gpurun_out
T=gpurun_out
run() {  # n tag extra...
  local n=$1 tag=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 --no-parity --no-reference-gpu "$@" > $T/r2q_bench_n${n}_${tag}.json 2> $T/r2q_bench_n${n}_${tag}.err
  echo "n=$n $tag rc=$?"; tail -2 $T/r2q_bench_n${n}_${tag}.err | cut -c1-200
  python - <<PY
import json
try:
    d = json.loads(open('$T/r2q_bench_n${n}_${tag}.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'comm_exposed_ms', 'gpu_launches')}, d.get('e2e'), d.get('clocks'))
except Exception as e:
    print('no line', e)
PY
}
run 8 same
run 8 perrank --per-rank-masks
run 4 same
run 2 same
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
  bench.py --gpus 8 --steps 2 --warmup 3 --no-parity --no-reference-gpu 2>&1 | grep -i "nvls\|channels\|P2P/CUMEM" | sort | uniq -c | sort -rn | head -8 > $T/r2q_nccl_info.txt
cat $T/r2q_nccl_info.txt | cut -c1-200
nvidia-smi topo -m > $T/r2q_topo.txt 2>&1; head -12 $T/r2q_topo.txt | cut -c1-200
