#!/bin/bash
# round 2, call zi (8 GPUs): the final build through the driver's own 8-GPU command line
mkdir -p gpurun_out
T=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 \
  bench.py --gpus 8 --steps 10 --warmup 3 > $T/r2zi_bench_n8.json 2> $T/r2zi_bench_n8.err
echo "n8 rc=$?"; grep "\[bench\]" $T/r2zi_bench_n8.err | head -2; tail -2 $T/r2zi_bench_n8.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zi_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['comm_exposed_ms'], d['e2e']['value'], d['clocks'])
PY
