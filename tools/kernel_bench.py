"""Per-kernel microbenchmarks on the ViT-L/16 step shapes (CUDA events, L2 flushed between
iterations by rotating over buffers larger than the 126 MB L2).  Prints one JSON line per case:
achieved TFLOP/s or GB/s and the fraction of the measured peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from avjepa_b200 import _cabi, engine  # noqa: E402
from avjepa_b200._cabi import BF16, F32, GEMM_NN, GEMM_NT, GEMM_TN  # noqa: E402

DEV = 'cuda'
PEAKS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) \
    else dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0)


ITERS, WARM, WITH_LIBRARY = 20, 3, True     # tools/ncu_cases.py shrinks these for profiler captures


def timeit(fn, iters=None, warm=None):
    iters = ITERS if iters is None else iters
    warm = WARM if warm is None else warm
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_gemm(layout, M, N, K, epi='none', tag=''):
    mode = engine.MODE_BF16
    a_shape = (K, M) if layout == GEMM_TN else (M, K)
    b_shape = (N, K) if layout == GEMM_NT else (K, N)
    nbuf = max(2, int(300e6 // max(1, (M * K + N * K + M * N) * 2)) + 1)
    nbuf = min(nbuf, 8)
    As = [torch.randn(a_shape, device=DEV).bfloat16() for _ in range(nbuf)]
    Bs = [torch.randn(b_shape, device=DEV).bfloat16() for _ in range(nbuf)]
    out_dtype = F32 if epi in ('accum', 'res') else BF16
    Cs = [torch.zeros((M, N), device=DEV, dtype=torch.float32 if out_dtype == F32 else torch.bfloat16) for _ in range(nbuf)]
    bias = torch.randn(N, device=DEV)
    res = torch.randn((M, N), device=DEV) if epi == 'res' else None
    pre = torch.empty((M, N), device=DEV, dtype=torch.bfloat16) if epi == 'gelu' else None
    i = [0]

    def fn():
        j = i[0] % nbuf
        i[0] += 1
        kw = {}
        if epi in ('bias', 'gelu', 'res'):
            kw['bias'] = bias.data_ptr()
        if epi == 'gelu':
            kw.update(act=1, pre_out=pre.data_ptr())
        if epi == 'res':
            kw['residual'] = res.data_ptr()
        if epi == 'accum':
            kw['accumulate'] = 1
        if epi == 'dact':
            kw['dact_aux'] = Cs[(j + 1) % nbuf].data_ptr()
        engine.gemm(mode, layout, As[j].data_ptr(), Bs[j].data_ptr(), Cs[j].data_ptr(), M, N, K, a_shape[1], b_shape[1], N,
                    out_dtype, **kw)
    ms = timeit(fn)
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    print(json.dumps(dict(kernel='gemm_umma', tag=tag, layout=['NT', 'NN', 'TN'][layout], M=M, N=N, K=K, epi=epi, ms=round(ms, 4),
                          tflops=round(tf, 1), frac_burst=round(tf / PEAKS['bf16_tflops'], 3))), flush=True)
    if not WITH_LIBRARY:
        return
    # cuBLAS (library) on the same shape, for context only
    A, B_ = As[0], Bs[0]
    Af = A.t() if layout == GEMM_TN else A
    Bf = B_.t() if layout == GEMM_NT else B_
    ms2 = timeit(lambda: torch.matmul(Af, Bf))
    print(json.dumps(dict(kernel='cublas(torch.matmul)', tag=tag, ms=round(ms2, 4),
                          tflops=round(2.0 * M * N * K / (ms2 * 1e-3) / 1e12, 1))), flush=True)


def bench_attn(B, N, H, hd, tag=''):
    qkv = torch.randn((B, N, 3, H, hd), device=DEV).bfloat16()
    out = torch.empty((B, N, H * hd), device=DEV, dtype=torch.bfloat16)
    lse = torch.empty((B, H, N), device=DEV)
    dout = torch.randn((B, N, H * hd), device=DEV).bfloat16()
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(int(_cabi.load().avj_attention_bwd_ws_floats(B, N, H, hd)), device=DEV)
    scale = hd ** -0.5
    s = engine.stream
    ms_f = timeit(lambda: _cabi.call('avj_attention_fwd', BF16, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd, scale, s()))
    ms_b = timeit(lambda: _cabi.call('avj_attention_bwd', BF16, qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                     dqkv.data_ptr(), ws.data_ptr(), B, N, H, hd, scale, s()))
    ff = 4.0 * B * H * N * N * hd
    print(json.dumps(dict(kernel='fa_fwd', tag=tag, B=B, N=N, H=H, hd=hd, ms=round(ms_f, 4), tflops=round(ff / ms_f / 1e9, 1),
                          frac_burst=round(ff / ms_f / 1e9 / PEAKS['bf16_tflops'], 3))), flush=True)
    print(json.dumps(dict(kernel='fa_bwd(dq+dkv)', tag=tag, ms=round(ms_b, 4), tflops_alg10=round(2.5 * ff / ms_b / 1e9, 1),
                          tflops_exec14=round(3.5 * ff / ms_b / 1e9, 1))), flush=True)
    if not WITH_LIBRARY:
        return
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    print(json.dumps(dict(kernel='torch SDPA fwd (library)', tag=tag, ms=round(ms_t, 4), tflops=round(ff / ms_t / 1e9, 1))), flush=True)


def bench_ln(rows, D):
    x = torch.randn((rows, D), device=DEV)
    g = torch.randn(D, device=DEV)
    b = torch.randn(D, device=DEV)
    y = torch.empty((rows, D), device=DEV, dtype=torch.bfloat16)
    mean = torch.empty(rows, device=DEV)
    rstd = torch.empty(rows, device=DEV)
    ms = timeit(lambda: engine.layernorm_fwd(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), BF16, mean.data_ptr(),
                                             rstd.data_ptr(), rows, D, 1e-6))
    gbs = rows * D * 6 / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='layernorm_fwd', rows=rows, D=D, ms=round(ms, 4), gbs=round(gbs, 1),
                          frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)
    dy = torch.randn((rows, D), device=DEV).bfloat16()
    dres = torch.randn((rows, D), device=DEV)
    dx = torch.empty((rows, D), device=DEV)
    dxl = torch.empty((rows, D), device=DEV, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    ws = torch.empty(int(_cabi.load().avj_layernorm_bwd_ws_floats(rows, D)), device=DEV)
    ms = timeit(lambda: engine.layernorm_bwd(dy.data_ptr(), BF16, x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             dres.data_ptr(), dx.data_ptr(), dxl.data_ptr(), BF16, dg.data_ptr(), db.data_ptr(),
                                             ws.data_ptr(), rows, D))
    gbs = rows * D * (2 + 4 + 4 + 4 + 2) / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='layernorm_bwd', rows=rows, D=D, ms=round(ms, 4), gbs=round(gbs, 1),
                          frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)


def bench_adamw(n):
    import ctypes as C
    p, g, m, v, k = (torch.randn(n, device=DEV) for _ in range(5))
    v.abs_()
    lp = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    klp = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    a = _cabi.AdamWArgs()
    a.p, a.g, a.m, a.v, a.target, a.p_lp, a.target_lp = (t.data_ptr() for t in (p, g, m, v, k, lp, klp))
    a.n = n
    a.lr, a.wd, a.beta1, a.beta2, a.eps, a.step, a.ema_m, a.skip_update, a.zero_grad = 1e-3, 0.05, 0.9, 0.999, 1e-8, 3, 0.998, 0, 1
    ms = timeit(lambda: _cabi.call('avj_adamw_ema_step', C.byref(a), engine.stream()))
    gbs = n * (16 + 12 + 4 + 8 + 4) / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='adamw_ema', n=n, ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)


def bench_gather(B, N, K, D):
    x = torch.randn((B, N, D), device=DEV)
    idx = torch.stack([torch.randperm(N, device=DEV)[:K].sort().values for _ in range(B)])
    out = torch.empty((B, K, D), device=DEV)
    ms = timeit(lambda: _cabi.call('avj_gather_rows_fwd', F32, x.data_ptr(), idx.data_ptr(), out.data_ptr(), B, N, K, D, engine.stream()))
    gbs = (2 * B * K * D * 4 + 8 * B * K) / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='gather_rows', B=B, N=N, K=K, D=D, ms=round(ms, 4), gbs=round(gbs, 1),
                          frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)


def bench_patch_embed(B, D, K=None):
    """im2col-free patch embedding of a [B, 3, 16, 224, 224] clip batch: every token (K None) or K kept tokens per clip."""
    x = torch.randn((B, 3, 16, 224, 224), device=DEV)
    w = torch.randn((D, 1536), device=DEV) * 0.02
    bias = torch.randn(D, device=DEV)
    pos = torch.randn((1568, D), device=DEV)
    Kt = K or 1568
    idx = torch.stack([torch.randperm(1568, device=DEV)[:Kt].sort().values for _ in range(B)]) if K else None
    out = torch.empty((B * Kt, D), device=DEV)
    ms = timeit(lambda: engine.patch_embed(x.data_ptr(), idx.data_ptr() if K else None, w.data_ptr(), out.data_ptr(), B, 3, 16, 224, 224, 2, 16,
                                           Kt, D, D, bias=bias.data_ptr(), pos=pos.data_ptr(), pos_idx=idx.data_ptr() if K else None,
                                           pos_rows=1568))
    fl = 2.0 * B * Kt * D * 1536
    print(json.dumps(dict(kernel='patch_embed (im2col-free, tf32)', B=B, tokens=B * Kt, D=D, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1),
                          clip_gbs=round(B * Kt * 1536 * 4 / (ms * 1e-3) / 1e9, 1))), flush=True)


def bench_loss(rows, D):
    z = torch.randn((rows, D), device=DEV)
    h = torch.randn((rows, D), device=DEV)
    dz = torch.empty_like(z)
    out = torch.zeros(1, device=DEV)
    ws = torch.empty(int(_cabi.load().avj_loss_ws_floats(z.numel())), device=DEV)
    ms = timeit(lambda: _cabi.call('avj_loss_fwd_bwd', z.data_ptr(), h.data_ptr(), dz.data_ptr(), out.data_ptr(), z.numel(), 2, 1.0, 0, 0.0,
                                   1.0, ws.data_ptr(), engine.stream()))
    gbs = rows * D * 12 / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='loss_fwd_bwd', rows=rows, D=D, ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)


def bench_colsum2(rows, D1, D2):
    a = torch.randn((rows, D1), device=DEV).bfloat16()
    b = torch.randn((rows, D2), device=DEV).bfloat16()
    o1 = torch.zeros(D1, device=DEV)
    o2 = torch.zeros(D2, device=DEV)
    ws = torch.empty(int(_cabi.load().avj_colsum_ws_floats(rows, D1 + D2)), device=DEV)
    ms = timeit(lambda: _cabi.call('avj_colsum2', a.data_ptr(), D1, D1, o1.data_ptr(), b.data_ptr(), D2, D2, o2.data_ptr(), BF16, rows,
                                   ws.data_ptr(), engine.stream()))
    gbs = rows * (D1 + D2) * 2 / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel='colsum2', rows=rows, D1=D1, D2=D2, ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / PEAKS['hbm_gbs'], 3))), flush=True)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    R_T, R_C, R_P = 24 * 1664, 24 * 384, 24 * 1216      # target / context / predictor rows at ViT-L, B=24
    if which in ('all', 'gemm'):
        for tag, M in (('target', R_T), ('ctx', R_C)):
            bench_gemm(GEMM_NT, M, 3072, 1024, 'bias', tag + ' qkv')
            bench_gemm(GEMM_NT, M, 1024, 1024, 'res', tag + ' proj')
            bench_gemm(GEMM_NT, M, 4096, 1024, 'gelu', tag + ' fc1')
            bench_gemm(GEMM_NT, M, 1024, 4096, 'res', tag + ' fc2')
        bench_gemm(GEMM_NN, R_C, 4096, 1024, 'dact', 'ctx fc2 dgrad')
        bench_gemm(GEMM_NN, R_C, 1024, 4096, 'none', 'ctx fc1 dgrad')
        bench_gemm(GEMM_TN, 4096, 1024, R_C, 'accum', 'ctx fc1 wgrad')
        bench_gemm(GEMM_TN, 1024, 4096, R_C, 'accum', 'ctx fc2 wgrad')
        bench_gemm(GEMM_NT, R_P, 1152, 384, 'bias', 'pred qkv')
        bench_gemm(GEMM_NT, R_P, 1536, 384, 'gelu', 'pred fc1')
        bench_gemm(GEMM_NT, R_P, 384, 1536, 'res', 'pred fc2')
        bench_gemm(GEMM_TN, 1536, 384, R_P, 'accum', 'pred fc1 wgrad')
        bench_gemm(GEMM_NN, R_P, 384, 1536, 'none', 'pred fc1 dgrad')
        bench_gemm(GEMM_NT, 8192, 8192, 8192, 'none', 'square 8192')
    if which in ('all', 'attn'):
        bench_attn(24, 1664, 16, 64, 'target enc')
        bench_attn(24, 384, 16, 64, 'ctx enc')
        bench_attn(24, 1216, 16, 24, 'predictor')
    if which == 'patch':
        bench_patch_embed(24, 1024)
        bench_patch_embed(24, 1024, K=184)
    if which == 'attn_h':                               # ViT-H (config 4): 16 heads of 80
        bench_attn(24, 1664, 16, 80, 'vit_huge target enc')
        bench_attn(24, 384, 16, 80, 'vit_huge ctx enc')
        bench_attn(8, 1664, 16, 128, 'hd128')
    if which in ('all', 'misc'):
        bench_ln(R_T, 1024)
        bench_ln(24 * 537, 1024)
        bench_ln(24 * 2450, 384)
        bench_colsum2(24 * 537, 3072, 4096)
        bench_colsum2(24 * 2450, 1152, 1536)
        bench_adamw(300_000_000)
        bench_gather(24, 1568, 800, 1024)
        bench_loss(24 * 1144, 1024)
