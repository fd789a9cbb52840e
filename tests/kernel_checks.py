"""Per-kernel numerics checks (GPU).  Each check calls ONE C-ABI entry point through ctypes and
compares it with a plain PyTorch fp32 computation of the same op on the same inputs.
Used by tests/test_kernels_gpu.py and by tools/gpu_bringup.py."""
import ctypes as C
import math

import torch

from avjepa_b200 import _cabi, engine
from avjepa_b200._cabi import BF16, F32, GEMM_NN, GEMM_NT, GEMM_TN, IDENTITY, RowMap

DEV = 'cuda'


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _tdt(code):
    return torch.bfloat16 if code == BF16 else torch.float32


def _tol(code):
    return 2e-2 if code == BF16 else 2e-5


def check_gemm(code, layout, M, N, K, epi='none', seed=0):
    """epi in none|bias|bias_gelu|bias_res|dact|accum|pos_map"""
    g = torch.Generator(device=DEV).manual_seed(seed)
    dt = _tdt(code)
    a_shape = (K, M) if layout == GEMM_TN else (M, K)
    b_shape = (N, K) if layout == GEMM_NT else (K, N)
    A = (torch.randn(a_shape, generator=g, device=DEV) * 0.5).to(dt)
    B_ = (torch.randn(b_shape, generator=g, device=DEV) * 0.5).to(dt)
    Af = A.float().t() if layout == GEMM_TN else A.float()
    Bf = B_.float().t() if layout == GEMM_NT else B_.float()
    ref = Af @ Bf
    kw, out_dtype, ldc, rows_out = {}, code, N, M
    keep = []
    if epi in ('bias', 'bias_gelu', 'bias_res', 'pos_map'):
        bias = torch.randn(N, generator=g, device=DEV)
        keep.append(bias)
        kw['bias'] = bias.data_ptr()
        ref = ref + bias
    pre = None
    if epi == 'bias_gelu':
        pre = torch.empty((M, N), dtype=dt, device=DEV)
        kw.update(act=1, pre_out=pre.data_ptr())
        pre_ref = ref.clone()
        ref = torch.nn.functional.gelu(ref)
    if epi == 'bias_res':
        res = torch.randn((M, N), generator=g, device=DEV)
        keep.append(res)
        kw['residual'] = res.data_ptr()
        out_dtype = F32
        ref = ref + res
    if epi == 'dact':
        aux = torch.randn((M, N), generator=g, device=DEV).to(dt)
        keep.append(aux)
        kw['dact_aux'] = aux.data_ptr()
        x = aux.float()
        cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
        pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
        ref = ref * (cdf + x * pdf)
    Cbuf = None
    if epi == 'accum':
        out_dtype = F32
        Cbuf = torch.randn((M, N), generator=g, device=DEV)
        ref = ref + Cbuf
        kw['accumulate'] = 1
    if epi == 'pos_map':
        # rows land in a [groups, stride, N] layout at an offset, plus gathered positional rows
        out_dtype = F32
        G = 4 if M % 4 == 0 else 1
        rpg = M // G
        stride, off, npos = rpg + 5, 3, 37
        pos = torch.randn((npos, N), generator=g, device=DEV)
        idx = torch.randint(0, npos, (M,), generator=g, device=DEV, dtype=torch.int64)
        keep += [pos, idx]
        kw.update(pos=pos.data_ptr(), pos_idx=idx.data_ptr(), pos_rows=npos, out_map=RowMap(rpg, stride, off))
        ref = ref + pos[idx]
        rows_out = G * stride
        Cbuf = torch.zeros((rows_out, N), device=DEV)
        full = torch.zeros((rows_out, N), device=DEV)
        r = torch.arange(M, device=DEV)
        full[(r // rpg) * stride + r % rpg + off] = ref
        ref = full
    if Cbuf is None:
        Cbuf = torch.empty((rows_out, N), dtype=_tdt(out_dtype), device=DEV)
    lda, ldb = a_shape[1], b_shape[1]
    engine.gemm(engine.Mode(code), layout, A.data_ptr(), B_.data_ptr(), Cbuf.data_ptr(), M, N, K, lda, ldb, ldc, out_dtype, **kw)
    torch.cuda.synchronize()
    err = _rel(Cbuf.float(), ref)
    tol = _tol(code) if out_dtype == code or code == BF16 else _tol(code)
    ok = err < tol
    if pre is not None:
        e2 = _rel(pre.float(), pre_ref)
        ok = ok and e2 < tol
        err = max(err, e2)
    return ok, err


def check_layernorm(y_code, rows, D, affine=True, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn((rows, D), generator=g, device=DEV) * 2 + 0.5
    gam = torch.randn(D, generator=g, device=DEV) if affine else None
    bet = torch.randn(D, generator=g, device=DEV) if affine else None
    y = torch.empty((rows, D), dtype=_tdt(y_code), device=DEV)
    mean = torch.empty(rows, device=DEV)
    rstd = torch.empty(rows, device=DEV)
    engine.layernorm_fwd(x.data_ptr(), gam.data_ptr() if affine else None, bet.data_ptr() if affine else None,
                         y.data_ptr(), y_code, mean.data_ptr(), rstd.data_ptr(), rows, D, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (D,), gam, bet, 1e-6)
    e_f = _rel(y.float(), ref)
    # backward
    dy = torch.randn((rows, D), generator=g, device=DEV).to(_tdt(y_code))
    dres = torch.randn((rows, D), generator=g, device=DEV)
    dx = torch.empty((rows, D), device=DEV)
    dx_lp = torch.empty((rows, D), dtype=_tdt(y_code), device=DEV)
    dg = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    lib = _cabi.load()
    ws = torch.empty(int(lib.avj_layernorm_bwd_ws_floats(rows, D)), device=DEV)
    engine.layernorm_bwd(dy.data_ptr(), y_code, x.data_ptr(), gam.data_ptr() if affine else None, mean.data_ptr(),
                         rstd.data_ptr(), dres.data_ptr(), dx.data_ptr(), dx_lp.data_ptr(), y_code,
                         dg.data_ptr() if affine else None, db.data_ptr() if affine else None, ws.data_ptr(), rows, D)
    xr = x.clone().requires_grad_(True)
    gr = gam.clone().requires_grad_(True) if affine else None
    br = bet.clone().requires_grad_(True) if affine else None
    torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy.float())
    e_dx = _rel(dx, xr.grad + dres)
    e_lp = _rel(dx_lp.float(), xr.grad + dres)
    errs = [e_f, e_dx, e_lp]
    if affine:
        errs += [_rel(dg, gr.grad), _rel(db, br.grad)]
    tol = _tol(y_code)
    return all(e < tol for e in errs), max(errs)


def check_attention(code, B, N, H, hd, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    dt = _tdt(code)
    qkv = (torch.randn((B, N, 3, H, hd), generator=g, device=DEV)).to(dt)
    out = torch.empty((B, N, H * hd), dtype=dt, device=DEV)
    lse = torch.empty((B, H, N), device=DEV)
    scale = hd ** -0.5
    _cabi.call('avj_attention_fwd', code, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd, scale, engine.stream())
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, N, H * hd)
    e_o = _rel(out.float(), ref)
    e_l = _rel(lse, torch.logsumexp(s, -1))
    dout = torch.randn((B, N, H * hd), generator=g, device=DEV).to(dt)
    dqkv = torch.empty_like(qkv)
    lib = _cabi.load()
    ws = torch.empty(int(lib.avj_attention_bwd_ws_floats(B, N, H, hd)), device=DEV)
    _cabi.call('avj_attention_bwd', code, qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
               ws.data_ptr(), B, N, H, hd, scale, engine.stream())
    ref.backward(dout.float())
    e_g = _rel(dqkv.float(), qf.grad)
    tol = 3e-2 if code == BF16 else 5e-5
    global LAST_ATTENTION_ERRORS
    LAST_ATTENTION_ERRORS = dict(out=e_o, lse=e_l, dq=_rel(dqkv[:, :, 0].float(), qf.grad[:, :, 0]),
                                 dk=_rel(dqkv[:, :, 1].float(), qf.grad[:, :, 1]), dv=_rel(dqkv[:, :, 2].float(), qf.grad[:, :, 2]))
    ok_parts = all(LAST_ATTENTION_ERRORS[k] < tol for k in ('dq', 'dk', 'dv'))
    return (e_o < tol and e_l < 1e-3 and e_g < tol and ok_parts), max(e_o, e_g)


LAST_ATTENTION_ERRORS = {}


def check_gather(code, B=3, N=1568, K=300, D=192, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    dt = _tdt(code)
    x = torch.randn((B, N, D), generator=g, device=DEV).to(dt)
    idx = torch.stack([torch.randperm(N, generator=g, device=DEV)[:K].sort().values for _ in range(B)])
    from avjepa_b200.src.masks.utils import apply_masks
    xr = x.clone().requires_grad_(True)
    out = apply_masks(xr, [idx])
    ref = torch.gather(x, 1, idx.unsqueeze(-1).repeat(1, 1, D))
    exact = torch.equal(out, ref)
    dout = torch.randn((B, K, D), generator=g, device=DEV).to(dt)
    out.backward(dout)
    dref = torch.zeros_like(x).scatter_add_(1, idx.unsqueeze(-1).repeat(1, 1, D), dout)
    exact_b = torch.equal(xr.grad, dref)
    return exact and exact_b, 0.0


def check_patchify(code, B=2, seed=0, audio=False):
    g = torch.Generator(device=DEV).manual_seed(seed)
    dt = _tdt(code)
    if audio:
        x = torch.randn((B, 1, 1, 128, 192), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 1, 1, 128, 192, 1, 96
        w = torch.randn((16, 1, 1, 16, 16), generator=g, device=DEV)
    else:
        x = torch.randn((B, 3, 16, 224, 224), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 3, 16, 224, 224, 2, 1568
        w = torch.randn((16, 3, 2, 16, 16), generator=g, device=DEV)
    K = 50
    idx = torch.stack([torch.randperm(ntok, generator=g, device=DEV)[:K].sort().values for _ in range(B)])
    kd = Cc * tub * 256
    out = torch.empty((B * K, kd), dtype=dt, device=DEV)
    _cabi.call('avj_patchify', x.data_ptr(), idx.data_ptr(), out.data_ptr(), code, B, Cc, T, H, W, tub, 16, K, engine.stream())
    xin = x.to(dt).float()
    conv = torch.nn.functional.conv3d(xin, w, stride=(tub, 16, 16)).flatten(2).transpose(1, 2)      # [B, ntok, 16]
    ref = torch.gather(conv, 1, idx.unsqueeze(-1).repeat(1, 1, 16)).reshape(B * K, 16)
    got = out.float() @ w.reshape(16, kd).t()
    e = _rel(got, ref)
    # full (no index) variant
    out2 = torch.empty((B * ntok, kd), dtype=dt, device=DEV)
    _cabi.call('avj_patchify', x.data_ptr(), None, out2.data_ptr(), code, B, Cc, T, H, W, tub, 16, ntok, engine.stream())
    e2 = _rel(out2.float() @ w.reshape(16, kd).t(), conv.reshape(B * ntok, 16))
    return (e < 1e-4 and e2 < 1e-4), max(e, e2)


def check_patch_embed(B=3, D=192, seed=0, audio=False, masked=True):
    """avj_patch_embed (im2col-free, tf32 tensor cores) against Conv3d + bias + positional rows in fp32; rows land in a
    wider sequence through the output row map like the encoder's [video | audio] layout."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    if audio:
        x = torch.randn((B, 1, 1, 128, 192), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 1, 1, 128, 192, 1, 96
    else:
        x = torch.randn((B, 3, 16, 224, 224), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 3, 16, 224, 224, 2, 1568
    w = torch.randn((D, Cc, tub, 16, 16), generator=g, device=DEV) * 0.05
    bias = torch.randn(D, generator=g, device=DEV)
    pos = torch.randn((ntok, D), generator=g, device=DEV)
    K = 77 if masked else ntok
    idx = torch.stack([torch.randperm(ntok, generator=g, device=DEV)[:K].sort().values for _ in range(B)]) if masked else None
    N, off = K + 5, 3                                            # each sample's rows sit at [off, off + K) of N
    out = torch.full((B * N, D), -7.0, device=DEV)
    engine.patch_embed(x.data_ptr(), idx.data_ptr() if masked else None, w.data_ptr(), out.data_ptr(), B, Cc, T, H, W, tub, 16, K, D, D,
                       bias=bias.data_ptr(), pos=pos.data_ptr(), pos_idx=idx.data_ptr() if masked else None, pos_rows=ntok,
                       out_map=RowMap(K, N, off))
    conv = (torch.nn.functional.conv3d(x.double(), w.double(), bias.double(), stride=(tub, 16, 16)).flatten(2).transpose(1, 2)
            + pos.double()).float()                                                                     # [B, ntok, D]
    ref = torch.gather(conv, 1, idx.unsqueeze(-1).repeat(1, 1, D)) if masked else conv
    got = out.view(B, N, D)
    e = _rel(got[:, off:off + K], ref)
    untouched = bool((got[:, :off] == -7.0).all()) and bool((got[:, off + K:] == -7.0).all())
    return (e < 2e-3 and untouched), e


def check_patch_embed_wgrad(B=3, D=192, seed=0, audio=False, masked=True):
    """avj_patch_embed_wgrad (patch values gathered by the GEMM producer, bf16 tensor cores, fp32 accumulate INTO gw) against
    the autograd weight gradient of Conv3d in fp64."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    if audio:
        x = torch.randn((B, 1, 1, 128, 192), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 1, 1, 128, 192, 1, 96
    else:
        x = torch.randn((B, 3, 16, 224, 224), generator=g, device=DEV)
        Cc, T, H, W, tub, ntok = 3, 16, 224, 224, 2, 1568
    kd = Cc * tub * 256
    K = 77 if masked else ntok
    idx = torch.stack([torch.randperm(ntok, generator=g, device=DEV)[:K].sort().values for _ in range(B)]) if masked else None
    dy = (torch.randn((B * K, D), generator=g, device=DEV) * 0.3).to(torch.bfloat16)
    gw0 = torch.randn((D, kd), generator=g, device=DEV)
    gw = gw0.clone()
    engine.patch_embed_wgrad(x.data_ptr(), idx.data_ptr() if masked else None, dy.data_ptr(), gw.data_ptr(), B, Cc, T, H, W, tub, 16, K, D)
    w = torch.zeros((D, Cc, tub, 16, 16), dtype=torch.float64, device=DEV, requires_grad=True)
    conv = torch.nn.functional.conv3d(x.to(torch.bfloat16).double(), w, stride=(tub, 16, 16)).flatten(2).transpose(1, 2)     # [B, ntok, D]
    sel = torch.gather(conv, 1, idx.unsqueeze(-1).repeat(1, 1, D)) if masked else conv
    (sel.reshape(B * K, D) * dy.double()).sum().backward()
    ref = gw0.double() + w.grad.reshape(D, kd)
    e = _rel(gw - gw0, ref - gw0.double())
    return e < 2e-3, e


def check_rows(seed=0):
    """copy_rows / colsum / fill_mask_tokens with non-trivial row maps."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    rows, D, rpg, stride, off = 24, 64, 6, 11, 2
    src = torch.randn((50, D), generator=g, device=DEV)
    dst = torch.zeros((rows, D), dtype=torch.bfloat16, device=DEV)
    engine.copy_rows(src.data_ptr(), F32, D, RowMap(rpg, stride, off), dst.data_ptr(), BF16, D, IDENTITY, rows, D)
    r = torch.arange(rows, device=DEV)
    phys = (r // rpg) * stride + r % rpg + off
    ok1 = torch.equal(dst, src[phys].to(torch.bfloat16))
    out = torch.zeros(D, device=DEV)
    lib = _cabi.load()
    ws = torch.empty(int(lib.avj_colsum_ws_floats(rows, D)), device=DEV)
    engine.colsum(src.data_ptr(), F32, D, RowMap(rpg, stride, off), out.data_ptr(), rows, D, ws.data_ptr())
    e2 = _rel(out, src[phys].sum(0))
    big = torch.randn((1000, 384), generator=g, device=DEV).to(torch.bfloat16)
    out3 = torch.ones(384, device=DEV)
    ws = torch.empty(int(lib.avj_colsum_ws_floats(1000, 384)), device=DEV)
    engine.colsum(big.data_ptr(), BF16, 384, IDENTITY, out3.data_ptr(), 1000, 384, ws.data_ptr())
    e3 = _rel(out3, big.float().sum(0) + 1)
    tok = torch.randn(D, generator=g, device=DEV)
    pos = torch.randn((30, D), generator=g, device=DEV)
    idx = torch.randint(0, 30, (rows,), generator=g, device=DEV, dtype=torch.int64)
    x = torch.zeros((50, D), device=DEV)
    _cabi.call('avj_fill_mask_tokens', tok.data_ptr(), pos.data_ptr(), idx.data_ptr(), x.data_ptr(), D,
               RowMap(rpg, stride, off), rows, D, engine.stream())
    ok4 = torch.equal(x[phys], tok + pos[idx])
    return ok1 and e2 < 1e-5 and e3 < 1e-5 and ok4, max(e2, e3)


def check_loss(seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    from avjepa_b200 import loss as L
    z = [torch.randn((2, 100, 192), generator=g, device=DEV).requires_grad_(True),
         torch.randn((2, 300, 192), generator=g, device=DEV).requires_grad_(True)]
    h = [torch.randn((2, 100, 192), generator=g, device=DEV), torch.randn((2, 300, 192), generator=g, device=DEV)]
    errs = []
    for p in (1.0, 2.0):
        for t in z:
            t.grad = None
        lo = L.jepa_loss(z, h, p)
        (lo * 3.0).backward()
        zr = [t.detach().clone().requires_grad_(True) for t in z]
        ref = sum(torch.mean(torch.abs(a - b) ** p) / p for a, b in zip(zr, h)) / 2
        (ref * 3.0).backward()
        errs += [abs(float(lo) - float(ref)) / abs(float(ref))] + [_rel(a.grad, b.grad) for a, b in zip(z, zr)]
    zz = [torch.randn((2, 64, 192), generator=g, device=DEV) * 0.3, torch.randn((2, 64, 192), generator=g, device=DEV)]
    rv = L.reg_value(zz)
    rr = L.reg_loss_differentiable(zz)
    errs.append(abs(float(rv) - float(rr)) / max(abs(float(rr)), 1e-6))
    return all(e < 1e-5 for e in errs), max(errs)


def check_adamw(seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    n = 4096 + 8
    p = torch.randn(n, generator=g, device=DEV)
    k = p.clone() + 0.1
    q = torch.nn.Parameter(p.clone())
    kr = k.clone()
    opt = torch.optim.AdamW([q], lr=3e-3, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    lp = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    klp = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    scale = torch.full((1,), 0.5, device=DEV)
    errs = []
    for step in range(1, 5):
        gr = torch.randn(n, generator=g, device=DEV)
        q.grad = gr.clone() * 0.5
        opt.step()
        kr.mul_(0.99).add_(0.01 * q.data)
        gbuf = gr.clone()
        a = _cabi.AdamWArgs()
        a.p, a.g, a.m, a.v = p.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr()
        a.target, a.p_lp, a.target_lp = k.data_ptr(), lp.data_ptr(), klp.data_ptr()
        a.n = n
        a.lr, a.wd, a.beta1, a.beta2, a.eps = 3e-3, 0.1, 0.9, 0.999, 1e-8
        a.step, a.ema_m, a.skip_update, a.zero_grad = step, 0.99, 0, 1
        a.scale_ptr = scale.data_ptr()
        _cabi.call('avj_adamw_ema_step', C.byref(a), engine.stream())
        errs += [_rel(p, q.data), _rel(k, kr), float(gbuf.abs().max())]
    errs.append(_rel(lp.float(), p.to(torch.bfloat16).float()))
    errs.append(_rel(klp.float(), k.to(torch.bfloat16).float()))
    x = torch.randn(100000, generator=g, device=DEV)
    out = torch.empty(1, device=DEV)
    lib = _cabi.load()
    ws = torch.empty(int(lib.avj_sumsq_ws_floats(x.numel())), device=DEV)
    _cabi.call('avj_sumsq', x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), engine.stream())
    errs.append(abs(float(out) - float((x.double() ** 2).sum())) / float((x.double() ** 2).sum()))
    coef = torch.empty(1, device=DEV)
    _cabi.call('avj_clip_coef', out.data_ptr(), 10.0, 1.0, coef.data_ptr(), engine.stream())
    errs.append(abs(float(coef) - min(1.0, 10.0 / (float(x.norm()) + 1e-6))))
    return all(e < 2e-6 for e in errs), max(errs)
