#!/bin/bash
# round 2, call s: dynamic unit scheduler of the persistent GEMMs: parity, per-shape times vs the static schedule, step
mkdir -p gpurun_out
T=gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "gemm" > $T/r2s_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $T/r2s_pytest.log | cut -c1-300
timeout 300 python tools/kernel_bench.py gemm > $T/r2s_gemm_dyn.log 2>&1
AVJ_GEMM_DYNAMIC=0 timeout 300 python tools/kernel_bench.py gemm > $T/r2s_gemm_static.log 2>&1
python - <<'PY'
import json
def load(f):
    out = {}
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if d.get('kernel', '').startswith('gemm_umma') or 'tag' in d and 'cublas' not in d.get('kernel', ''):
            out[d.get('tag')] = d
    return out
a, b = load('gpurun_out/r2s_gemm_dyn.log'), load('gpurun_out/r2s_gemm_static.log')
for k in a:
    if k in b: print(f"{k:24s} dyn {a[k]['ms']:.4f} ms {a[k].get('tflops')}  static {b[k]['ms']:.4f} ms {b[k].get('tflops')}")
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity > $T/r2s_bench.json 2> $T/r2s_bench.err
echo "bench rc=$?"; tail -2 $T/r2s_bench.err | cut -c1-300; cut -c1-400 $T/r2s_bench.json
AVJ_GEMM_DYNAMIC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-reference-gpu --no-parity > $T/r2s_bench_static.json 2> $T/r2s_bench_static.err
echo "bench static rc=$?"; cut -c1-400 $T/r2s_bench_static.json
true

