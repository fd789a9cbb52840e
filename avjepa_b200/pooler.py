"""Forward/backward schedule of the attentive pooler's cross-attention block (frozen-eval consumer).

Restates ``CrossAttentionBlock.forward`` / ``CrossAttention.forward`` of the reference
(``src/models/utils/modules.py:179-183``, ``:139-159``) as an explicit kernel schedule, one autograd node per call:

    h  = LN1(x)                      x: [B, N, D] encoder tokens (LN1 only in the complete block)
    kv = h Wkv^T + bkv               one tensor-core GEMM over all B*N tokens
    qp = q Wq^T + bq                 q: [B, n, D] learned queries (n = 1 in the reference's classifiers)
    o  = softmax(qp k^T / sqrt(hd)) v   avj_xattn_fwd: every K / V row read once, HBM-bound
    q1 = q + o Wproj^T + bproj
    q2 = q1 + fc2(gelu(fc1(LN2(q1))))   (complete block only)

The backward accumulates parameter gradients in place (like the backbone nodes) and returns gradients for ``q`` and --
only when it requires grad, which a frozen encoder's output does not -- for ``x``.
"""
import torch

from avjepa_b200 import _cabi, engine
from avjepa_b200._cabi import F32, GEMM_NN, GEMM_NT, GEMM_TN, IDENTITY
from avjepa_b200.backbone import _f32c, _shadows
from avjepa_b200.engine import LinearW, NormW, rowmap, stream


class _St(object):
    pass


def _forward(xattn, block, q, x, save, mode):
    engine.require_cuda(x, 'pooler tokens')
    dev = x.device
    B, N, D = x.shape
    n = q.shape[1]
    H = xattn.num_heads
    hd = D // H
    cd, s = mode.code, mode.size
    R, Q = B * N, B * n
    x = _f32c(x)
    qin = _f32c(q.expand(B, n, D) if q.shape[0] != B else q)
    sh = _shadows(xattn)
    st = _St()
    st.dims = (B, N, n, D, H, hd)
    st.mode, st.x, st.qin = mode, x, qin
    al = engine._align
    hid = block.mlp.fc1.out_features if block is not None else 0
    arena = engine.Arena(al(R * D * s) + 2 * al(R * 4) + al(R * 2 * D * s) + 4 * al(Q * D * s) + al(B * H * n * 4) + 2 * al(Q * D * 4)
                         + 2 * al(Q * 4) + 2 * al(Q * max(hid, 1) * s) + 4096, dev)
    a = arena.alloc
    st.arena = arena
    # ---- token side: LN1 (complete block) or a cast, then the kv Linear
    st.h, st.mean1, st.rstd1 = a(R * D * s), a(R * 4), a(R * 4)
    if block is not None:
        n1 = NormW(block.norm1, False)
        engine.layernorm_fwd(x.data_ptr(), n1.w, n1.b, st.h, cd, st.mean1, st.rstd1, R, D, n1.eps)
    else:
        engine.copy_rows(x.data_ptr(), F32, D, IDENTITY, st.h, cd, D, IDENTITY, R, D)
    kvw = LinearW(xattn.kv, sh, mode, False)
    st.kv = a(R * 2 * D * s)
    engine.gemm(mode, GEMM_NT, st.h, kvw.w, st.kv, R, 2 * D, D, D, D, 2 * D, cd, bias=kvw.b)
    # ---- query side
    st.qc, st.qp, st.o, st.lse = a(Q * D * s), a(Q * D * s), a(Q * D * s), a(B * H * n * 4)
    engine.copy_rows(qin.data_ptr(), F32, D, IDENTITY, st.qc, cd, D, IDENTITY, Q, D)
    qw = LinearW(xattn.q, sh, mode, False)
    engine.gemm(mode, GEMM_NT, st.qc, qw.w, st.qp, Q, D, D, D, D, D, cd, bias=qw.b)
    _cabi.call('avj_xattn_fwd', cd, st.qp, st.kv, st.o, st.lse, B, N, n, H, hd, float(hd ** -0.5), stream())
    pw = LinearW(xattn.proj, sh, mode, False)
    if block is None:                                   # bare CrossAttention: just the projection
        out = torch.empty((B, n, D), dtype=torch.float32, device=dev)
        engine.gemm(mode, GEMM_NT, st.o, pw.w, out.data_ptr(), Q, D, D, D, D, D, F32, bias=pw.b)
        return out, st
    st.q1 = a(Q * D * 4)
    engine.gemm(mode, GEMM_NT, st.o, pw.w, st.q1, Q, D, D, D, D, D, F32, bias=pw.b, residual=qin.data_ptr())
    n2 = NormW(block.norm2, False)
    st.h2, st.mean2, st.rstd2 = a(Q * D * s), a(Q * 4), a(Q * 4)
    engine.layernorm_fwd(st.q1, n2.w, n2.b, st.h2, cd, st.mean2, st.rstd2, Q, D, n2.eps)
    f1, f2 = LinearW(block.mlp.fc1, sh, mode, False), LinearW(block.mlp.fc2, sh, mode, False)
    st.pre, st.act = a(Q * hid * s), a(Q * hid * s)
    engine.gemm(mode, GEMM_NT, st.h2, f1.w, st.act, Q, hid, D, D, D, hid, cd, bias=f1.b, act=1, pre_out=st.pre if save else None)
    out = torch.empty((B, n, D), dtype=torch.float32, device=dev)
    engine.gemm(mode, GEMM_NT, st.act, f2.w, out.data_ptr(), Q, D, hid, hid, hid, D, F32, bias=f2.b, residual=st.q1)
    return out, st


def _backward(xattn, block, st, dout, need_dx):
    B, N, n, D, H, hd = st.dims
    mode = st.mode
    cd, s = mode.code, mode.size
    R, Q = B * N, B * n
    dev = dout.device
    dout = _f32c(dout)
    sh = _shadows(xattn)
    lib = _cabi.load()
    al = engine._align
    hid = block.mlp.fc1.out_features if block is not None else 0
    ws_f = max(lib.avj_colsum_ws_floats(max(R, Q), max(2 * D, hid, D)), lib.avj_layernorm_bwd_ws_floats(max(R, Q), D))
    need = (4 * al(Q * D * 4) + 6 * al(Q * D * s) + 2 * al(Q * max(hid, 1) * s) + al(R * 2 * D * s) + al(R * D * s) + al(R * D * 4)
            + al(R * D * s) + 4 * ws_f + (1 << 16))
    sc = engine.SCRATCH.get(need, dev)
    ws = sc.alloc(4 * ws_f)
    d_lp = sc.alloc(Q * D * s)                   # compute-dtype copy of the current query-side gradient
    pw = LinearW(xattn.proj, sh, mode, True)
    if block is not None:
        f1, f2 = LinearW(block.mlp.fc1, sh, mode, True), LinearW(block.mlp.fc2, sh, mode, True)
        n2 = NormW(block.norm2, True)
        engine.copy_rows(dout.data_ptr(), F32, D, IDENTITY, d_lp, cd, D, IDENTITY, Q, D)
        if f2.gb:
            engine.colsum(dout.data_ptr(), F32, D, IDENTITY, f2.gb, Q, D, ws)
        if f2.gw:
            engine.gemm(mode, GEMM_TN, d_lp, st.act, f2.gw, D, hid, Q, D, hid, hid, F32, accumulate=1)
        d_hid = sc.alloc(Q * hid * s)
        engine.gemm(mode, GEMM_NN, d_lp, f2.w, d_hid, Q, hid, D, D, hid, hid, cd, dact_aux=st.pre)
        if f1.gb:
            engine.colsum(d_hid, cd, hid, IDENTITY, f1.gb, Q, hid, ws)
        if f1.gw:
            engine.gemm(mode, GEMM_TN, d_hid, st.h2, f1.gw, hid, D, Q, hid, D, D, F32, accumulate=1)
        d_h2 = sc.alloc(Q * D * s)
        engine.gemm(mode, GEMM_NN, d_hid, f1.w, d_h2, Q, D, hid, hid, D, D, cd)
        dq1 = sc.alloc(Q * D * 4)
        engine.layernorm_bwd(d_h2, cd, st.q1, n2.w, st.mean2, st.rstd2, dout.data_ptr(), dq1, d_lp, cd, n2.gw, n2.gb, ws, Q, D)
    else:
        dq1 = dout.data_ptr()
        engine.copy_rows(dq1, F32, D, IDENTITY, d_lp, cd, D, IDENTITY, Q, D)
    # ---- proj
    if pw.gb:
        engine.colsum(dq1, F32, D, IDENTITY, pw.gb, Q, D, ws)
    if pw.gw:
        engine.gemm(mode, GEMM_TN, d_lp, st.o, pw.gw, D, D, Q, D, D, D, F32, accumulate=1)
    d_o = sc.alloc(Q * D * s)
    engine.gemm(mode, GEMM_NN, d_lp, pw.w, d_o, Q, D, D, D, D, D, cd)
    # ---- cross-attention
    d_qp, d_kv = sc.alloc(Q * D * s), sc.alloc(R * 2 * D * s)
    _cabi.call('avj_xattn_bwd', cd, st.qp, st.kv, st.o, d_o, st.lse, d_qp, d_kv, B, N, n, H, hd, float(hd ** -0.5), stream())
    # ---- q Linear
    qw = LinearW(xattn.q, sh, mode, True)
    if qw.gb:
        engine.colsum(d_qp, cd, D, IDENTITY, qw.gb, Q, D, ws)
    if qw.gw:
        engine.gemm(mode, GEMM_TN, d_qp, st.qc, qw.gw, D, D, Q, D, D, D, F32, accumulate=1)
    dq = torch.empty((B, n, D), dtype=torch.float32, device=dev)
    # d q = d q1 (residual path of the complete block) + d_qp Wq
    engine.gemm(mode, GEMM_NN, d_qp, qw.w, dq.data_ptr(), Q, D, D, D, D, D, F32, residual=dq1 if block is not None else None)
    # ---- kv Linear and LN1
    kvw = LinearW(xattn.kv, sh, mode, True)
    if kvw.gb:
        engine.colsum(d_kv, cd, 2 * D, IDENTITY, kvw.gb, R, 2 * D, ws)
    if kvw.gw:
        engine.gemm(mode, GEMM_TN, d_kv, st.h, kvw.gw, 2 * D, D, R, 2 * D, D, D, F32, accumulate=1)
    dx = None
    n1_trainable = block is not None and (block.norm1.weight.requires_grad or block.norm1.bias.requires_grad)
    if need_dx or n1_trainable:
        d_h = sc.alloc(R * D * s)
        engine.gemm(mode, GEMM_NN, d_kv, kvw.w, d_h, R, D, 2 * D, 2 * D, D, D, cd)
        if block is not None:
            n1 = NormW(block.norm1, True)
            dxt = torch.empty((B, N, D), dtype=torch.float32, device=dev) if need_dx else None
            dx_ptr = dxt.data_ptr() if need_dx else sc.alloc(R * D * 4)
            engine.layernorm_bwd(d_h, cd, st.x.data_ptr(), n1.w, st.mean1, st.rstd1, None, dx_ptr, None, cd, n1.gw, n1.gb, ws, R, D)
            dx = dxt
        else:
            dx = torch.empty((B, N, D), dtype=torch.float32, device=dev)
            engine.copy_rows(d_h, cd, D, IDENTITY, dx.data_ptr(), F32, D, IDENTITY, R, D)
    return dq, dx


class CrossAttnFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, xattn, block, save, mode, q, x, *params):
        out, st = _forward(xattn, block, q, x, save, mode)
        ctx.args = (xattn, block, st if save else None, len(params), tuple(q.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        xattn, block, st, n_params, q_shape = ctx.args
        dq, dx = _backward(xattn, block, st, dout, ctx.needs_input_grad[5])
        if q_shape[0] != dq.shape[0]:                    # queries were broadcast over the batch
            dq = dq.sum(dim=0, keepdim=True)
        ctx.args = None
        return (None, None, None, None, dq, dx) + (None,) * n_params


def run_cross_attention(xattn, block, q, x):
    """xattn: CrossAttention module; block: the enclosing CrossAttentionBlock or None (bare cross-attention)."""
    owner = block if block is not None else xattn
    params = [p for p in owner.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and (len(params) > 0 or q.requires_grad or x.requires_grad)
    return CrossAttnFn.apply(xattn, block, save, engine.Mode.current(), q, x, *params)
