#!/bin/bash
# round 2, call zh (4 GPUs): the final build through the driver's own multi-GPU command line
mkdir -p gpurun_out
T=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 \
  bench.py --gpus 4 --steps 10 --warmup 3 > $T/r2zh_bench_n4.json 2> $T/r2zh_bench_n4.err
echo "n4 rc=$?"; grep "\[bench\]" $T/r2zh_bench_n4.err; tail -2 $T/r2zh_bench_n4.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zh_bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['comm_exposed_ms'], d['e2e']['value'], d['clocks'])
PY
