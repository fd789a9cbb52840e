"""Forward/backward schedules of the encoder and predictor backbones and their autograd nodes.

One ``torch.autograd.Function`` per backbone call: the forward runs patch/ctx embedding and the
whole transformer stack as raw kernel launches and keeps the activations in a
:class:`~avjepa_b200.engine.StackRun`; the backward replays the stack in reverse, ACCUMULATING
parameter gradients straight into ``param.grad`` (created on demand) and returning only the
gradients of tensor inputs.  Reference call sites restated here:

* encoder   -- ``src/models/audiovision_transformer.py:186-239``, ``vision_transformer.py:162-201``
* predictor -- ``src/models/audiovisionpredictor.py:202-301``, ``predictor.py:175-239``
"""
import torch

from avjepa_b200 import _cabi, engine
from avjepa_b200._cabi import F32, GEMM_NN, GEMM_NT, GEMM_TN, IDENTITY
from avjepa_b200.engine import BlockW, LinearW, NormW, StackRun, rowmap, stream


def _shadows(mod):
    sh = mod.__dict__.get('_avj_shadows')
    if sh is None:
        sh = engine.Shadows()
        mod.__dict__['_avj_shadows'] = sh      # plain attribute: not a buffer, not in state_dict
    return sh


def _idx(m, device):
    """[B, K] int64 contiguous index tensor on `device` (or None)."""
    if m is None:
        return None
    if isinstance(m, (list, tuple)):
        if len(m) != 1:
            raise NotImplementedError('one mask per backbone call (use the MultiMask wrappers for several)')
        m = m[0]
    if m.device != device or m.dtype != torch.int64 or not m.is_contiguous():
        m = m.to(device=device, dtype=torch.int64).contiguous()
    return m


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def _rows_view(t):
    """(ptr, ld, RowMap) reading a [B, K, D] fp32 tensor row by row without copying when its
    strides allow (e.g. ``z[:, :K]`` slices produced by torch.split), else a contiguous copy."""
    B, K, D = t.shape
    if t.dtype == torch.float32 and t.stride(2) == 1 and t.stride(1) == D and (B == 1 or t.stride(0) % D == 0):
        gs = t.stride(0) // D if B > 1 else K
        return t, t.data_ptr(), D, rowmap(K, gs, 0)
    t = _f32c(t)
    return t, t.data_ptr(), D, IDENTITY


def _ptr(t):
    return None if t is None else t.data_ptr()


def _wptr(conv_or_lin_weight, mod, mode):
    return _shadows(mod).weight_ptr(conv_or_lin_weight, mode)


# ================================================================================================
# encoder
# ================================================================================================
class _EncState(object):
    pass


def _pos_table(mod, x, name):
    pos = getattr(mod, name)
    if hasattr(mod, 'interpolate_pos_encoding') and name != 'audio_pos_embed':
        pos = mod.interpolate_pos_encoding(x, pos)
    return _f32c(pos)


def encoder_forward(mod, x, y, mv, ma, save, mode):
    """Returns (out [B, N, D] fp32, state).  `y`/`ma` are None for the video-only encoder."""
    engine.require_cuda(x, 'encoder input')
    dev = x.device
    x = _f32c(x)
    if x.dim() == 4:                     # image model: [B, C, H, W] == one frame, tubelet 1
        x5 = x.unsqueeze(2)
        tub = 1
    else:
        x5 = x
        tub = mod.tubelet_size
    B, Cin, T, H, W = x5.shape
    p, D = mod.patch_size, mod.embed_dim
    n_full_v = (T // tub) * (H // p) * (W // p)
    mv, ma = _idx(mv, dev), _idx(ma, dev)
    Kv = mv.shape[1] if mv is not None else n_full_v
    Ka, n_full_a = 0, 0
    if y is not None:
        y = _f32c(y)
        n_full_a = (y.shape[2] // p) * (y.shape[3] // p)
        Ka = ma.shape[1] if ma is not None else n_full_a
    N = Kv + Ka
    cd = mode.code
    run = StackRun(B, N, D, mod.num_heads, len(mod.blocks), mode, save, dev)
    st = _EncState()
    st.run, st.B, st.N, st.Kv, st.Ka, st.mode = run, B, N, Kv, Ka, mode

    pos_name = 'video_pos_embed' if hasattr(mod, 'video_pos_embed') else 'pos_embed'
    pos_v = _pos_table(mod, x, pos_name)
    kd_v = Cin * tub * p * p
    st.patches_v = torch.empty((B * Kv, kd_v), dtype=mode.torch_dtype, device=dev)
    _cabi.call('avj_patchify', x5.data_ptr(), _ptr(mv), st.patches_v.data_ptr(), cd, B, Cin, T, H, W, tub, p, Kv, stream())
    pe = mod.patch_embed
    engine.gemm(mode, GEMM_NT, st.patches_v.data_ptr(), _wptr(pe.proj.weight, mod, mode), run.x_in(0),
                B * Kv, D, kd_v, kd_v, kd_v, D, F32, bias=_ptr(pe.proj.bias), pos=pos_v.data_ptr(),
                pos_idx=_ptr(mv), pos_rows=pos_v.shape[1], out_map=rowmap(Kv, N, 0))
    st.patches_a = None
    if y is not None and Ka > 0:
        pos_a = _f32c(mod.audio_pos_embed)
        kd_a = y.shape[1] * p * p
        st.patches_a = torch.empty((B * Ka, kd_a), dtype=mode.torch_dtype, device=dev)
        _cabi.call('avj_patchify', y.data_ptr(), _ptr(ma), st.patches_a.data_ptr(), cd, B, y.shape[1], 1, y.shape[2],
                   y.shape[3], 1, p, Ka, stream())
        engine.gemm(mode, GEMM_NT, st.patches_a.data_ptr(), _wptr(pe.audio_proj.weight, mod, mode), run.x_in(0),
                    B * Ka, D, kd_a, kd_a, kd_a, D, F32, bias=_ptr(pe.audio_proj.bias), pos=pos_a.data_ptr(),
                    pos_idx=_ptr(ma), pos_rows=pos_a.shape[1], out_map=rowmap(Ka, N, Kv))
    st.keep = (x5, y, mv, ma, pos_v)       # keep inputs alive until the kernels reading them retire
    sh = _shadows(mod)
    blocks = [BlockW(b, sh, mode, False) for b in mod.blocks]

    if mod.out_layers is not None:
        if save:
            raise NotImplementedError('out_layers is a frozen-eval feature; call under torch.no_grad()')
        outs = []
        norm = NormW(mod.norm, False)
        for i in range(len(blocks)):
            run.forward_layer(i, blocks[i])
            if i in mod.out_layers:
                o = torch.empty((B, N, D), dtype=torch.float32, device=dev)
                engine.layernorm_fwd(run.x_in(i + 1), norm.w, norm.b, o.data_ptr(), F32, None, None, B * N, D, norm.eps)
                outs.append(o)
        st.run_keepalive = run
        return outs, st

    out = torch.empty((B, N, D), dtype=torch.float32, device=dev)
    norm = NormW(mod.norm, False) if mod.norm is not None else None
    run.forward(blocks, norm, out.data_ptr(), F32)
    if not save:
        st.patches_v = st.patches_a = None
    return out, st


def encoder_backward(mod, st, dout):
    run, mode = st.run, st.mode
    B, N, Kv, Ka, D = st.B, st.N, st.Kv, st.Ka, run.D
    cd, s = mode.code, mode.size
    dev = dout.device
    dout = _f32c(dout)
    lib = _cabi.load()
    extra = engine._align(B * max(Kv, Ka, 1) * D * s) * 2 + 4 * lib.avj_colsum_ws_floats(B * max(Kv, Ka, 1), D) + (1 << 16)
    sc = engine.SCRATCH.get(run.scratch_bytes() + extra, dev)
    sh = _shadows(mod)
    blocks = [BlockW(b, sh, mode, True) for b in mod.blocks]
    norm = NormW(mod.norm, True) if mod.norm is not None else None
    from avjepa_b200 import dist as avj_dist
    sync = avj_dist.active_sync()
    evs = sync.layer_events_for(mod, run.L) if sync is not None else None
    dx0 = run.backward(blocks, norm, dout.data_ptr(), F32, sc, layer_events=evs)
    if evs is not None:
        sync.on_layers_enqueued(mod, evs)
    pe = mod.patch_embed
    ws = sc.alloc(4 * lib.avj_colsum_ws_floats(B * max(Kv, Ka, 1), D))
    dxc = sc.alloc(B * max(Kv, Ka, 1) * D * s)
    for (K, off, patches, conv) in ((Kv, 0, st.patches_v, pe.proj),
                                    (Ka, Kv, st.patches_a, getattr(pe, 'audio_proj', None))):
        if K == 0 or patches is None or conv is None:
            continue
        gw, gb = engine.grad_ptr(conv.weight), engine.grad_ptr(conv.bias) if conv.bias is not None else None
        kd = patches.shape[1]
        rm = rowmap(K, N, off)
        if gb is not None:
            engine.colsum(dx0, F32, D, rm, gb, B * K, D, ws)
        if gw is not None:
            engine.copy_rows(dx0, F32, D, rm, dxc, cd, D, IDENTITY, B * K, D)
            engine.gemm(mode, GEMM_TN, dxc, patches.data_ptr(), gw, D, kd, B * K, D, kd, kd, F32, accumulate=1)
    if sync is not None:
        sync.on_backward_done('encoder', mod)


class EncoderFn(torch.autograd.Function):
    """autograd node for one encoder call.  Inputs after `ma` are the module's parameters: they
    are listed so autograd knows the output depends on them; their gradients are accumulated
    in place by the backward kernels and `None` is returned for them."""

    @staticmethod
    def forward(ctx, mod, save, mode, x, y, mv, ma, *params):
        out, st = encoder_forward(mod, x, y, mv, ma, save, mode)
        ctx.mod, ctx.st, ctx.n_params = mod, st, len(params)
        if isinstance(out, list):
            return tuple(out)
        return out

    @staticmethod
    def backward(ctx, *douts):
        dout = douts[0]
        if dout is not None:
            encoder_backward(ctx.mod, ctx.st, dout)
        ctx.st = None
        return (None,) * (7 + ctx.n_params)


def run_encoder(mod, x, y, masks_v, masks_a):
    params = [p for p in mod.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and len(params) > 0
    out = EncoderFn.apply(mod, save, engine.Mode.current(), x, y, masks_v, masks_a, *params)
    if isinstance(out, tuple):
        return list(out)
    return out


# ================================================================================================
# standalone blocks (Block.forward)
# ================================================================================================
class _BlocksState(object):
    pass


class BlocksFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, owner, blocks_mod, norm_mod, save, mode, x, *params):
        engine.require_cuda(x, 'block input')
        B, N, D = x.shape
        x = _f32c(x)
        run = StackRun(B, N, D, blocks_mod[0].attn.num_heads, len(blocks_mod), mode, save, x.device,
                       hidden=blocks_mod[0].mlp.fc1.out_features)
        engine.copy_rows(x.data_ptr(), F32, D, IDENTITY, run.x_in(0), F32, D, IDENTITY, B * N, D)
        sh = _shadows(owner)
        out = torch.empty((B, N, D), dtype=torch.float32, device=x.device)
        run.forward([BlockW(b, sh, mode, False) for b in blocks_mod], NormW(norm_mod, False) if norm_mod is not None else None,
                    out.data_ptr(), F32)
        ctx.args = (owner, blocks_mod, norm_mod, run, mode, len(params))
        return out

    @staticmethod
    def backward(ctx, dout):
        owner, blocks_mod, norm_mod, run, mode, n_params = ctx.args
        dout = _f32c(dout)
        sc = engine.SCRATCH.get(run.scratch_bytes(), dout.device)
        sh = _shadows(owner)
        dx0 = run.backward([BlockW(b, sh, mode, True) for b in blocks_mod],
                           NormW(norm_mod, True) if norm_mod is not None else None, dout.data_ptr(), F32, sc)
        dx = torch.empty((run.B, run.N, run.D), dtype=torch.float32, device=dout.device)
        engine.copy_rows(dx0, F32, run.D, IDENTITY, dx.data_ptr(), F32, run.D, IDENTITY, run.R, run.D)
        return (None, None, None, None, None, dx) + (None,) * n_params


def run_blocks(owner, blocks_mod, norm_mod, x):
    params = [p for b in blocks_mod for p in b.parameters() if p.requires_grad]
    if norm_mod is not None:
        params += [p for p in norm_mod.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and (len(params) > 0 or x.requires_grad)
    return BlocksFn.apply(owner, blocks_mod, norm_mod, save, engine.Mode.current(), x, *params)


# ================================================================================================
# predictor
# ================================================================================================
class _PredState(object):
    pass


def predictor_forward(mod, parts, mask_index, z_v, z_a, mcv, mca, mtv, mta, save, mode):
    """parts = (embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a); the *_a entries are None for
    the video-only predictor.  Token layout [ctx_v | tgt_v | ctx_a | tgt_a].
    Returns (out [B, Ktv+Kta, D_enc] fp32, state)."""
    embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a = parts
    if tokens_v is None:
        raise NotImplementedError('predictor without mask tokens (the diffusion branch) is not implemented; '
                                  'every AV-JEPA config sets use_mask_tokens=True')
    engine.require_cuda(z_v, 'predictor context')
    dev = z_v.device
    B = z_v.shape[0]
    mcv, mtv, mca, mta = _idx(mcv, dev), _idx(mtv, dev), _idx(mca, dev), _idx(mta, dev)
    Kcv, Ktv = mcv.shape[1], mtv.shape[1]
    Kca = mca.shape[1] if (mca is not None and z_a is not None) else 0
    Kta = mta.shape[1] if mta is not None else 0
    if z_v.shape[1] != Kcv or (z_a is not None and z_a.shape[1] != Kca):
        raise ValueError(f'context token count {z_v.shape[1]} does not match the context mask ({Kcv})')
    N = Kcv + Ktv + Kca + Kta
    Kt = Ktv + Kta
    De, Dp = embed_v.in_features, embed_v.out_features
    cd, s = mode.code, mode.size
    mi = mask_index % len(tokens_v)
    run = StackRun(B, N, Dp, mod.predictor_blocks[0].attn.num_heads, len(mod.predictor_blocks), mode, save, dev)
    st = _PredState()
    st.run, st.B, st.N, st.mode, st.mi = run, B, N, mode, mi
    st.K = (Kcv, Ktv, Kca, Kta)
    st.masks = (mcv, mtv, mca, mta)
    sh = _shadows(mod)
    x0 = run.x_in(0)
    # ---- context rows: x = embed(z) + pos[idx]  (bias and gathered pos fused in the epilogue)
    keep, _, _, zmap_v = _rows_view(z_v)
    st.zc_v = torch.empty((B * Kcv, De), dtype=mode.torch_dtype, device=dev)
    engine.copy_rows(keep.data_ptr(), F32, De, zmap_v, st.zc_v.data_ptr(), cd, De, IDENTITY, B * Kcv, De)
    pv = _f32c(pos_v)
    engine.gemm(mode, GEMM_NT, st.zc_v.data_ptr(), sh.weight_ptr(embed_v.weight, mode), x0, B * Kcv, Dp, De, De, De, Dp, F32,
                bias=_ptr(embed_v.bias), pos=pv.data_ptr(), pos_idx=mcv.data_ptr(), pos_rows=pv.shape[1],
                out_map=rowmap(Kcv, N, 0))
    engine_keep = [keep, pv]
    st.zc_a = None
    pa = None
    if Kca > 0:
        keep_a, _, _, zmap_a = _rows_view(z_a)
        st.zc_a = torch.empty((B * Kca, De), dtype=mode.torch_dtype, device=dev)
        engine.copy_rows(keep_a.data_ptr(), F32, De, zmap_a, st.zc_a.data_ptr(), cd, De, IDENTITY, B * Kca, De)
        pa = _f32c(pos_a)
        engine.gemm(mode, GEMM_NT, st.zc_a.data_ptr(), sh.weight_ptr(embed_a.weight, mode), x0, B * Kca, Dp, De, De, De, Dp,
                    F32, bias=_ptr(embed_a.bias), pos=pa.data_ptr(), pos_idx=mca.data_ptr(), pos_rows=pa.shape[1],
                    out_map=rowmap(Kca, N, Kcv + Ktv))
        engine_keep += [keep_a, pa]
    elif Kta > 0:
        pa = _f32c(pos_a)
        engine_keep.append(pa)
    # ---- target rows: mask token + pos[idx]
    _cabi.call('avj_fill_mask_tokens', tokens_v[mi].data_ptr(), pv.data_ptr(), mtv.data_ptr(), x0, Dp,
               rowmap(Ktv, N, Kcv), B * Ktv, Dp, stream())
    if Kta > 0:
        _cabi.call('avj_fill_mask_tokens', tokens_a[mi].data_ptr(), pa.data_ptr(), mta.data_ptr(), x0, Dp,
                   rowmap(Kta, N, Kcv + Ktv + Kca), B * Kta, Dp, stream())
    st.keep = engine_keep
    # ---- blocks + predictor_norm, then project the target rows back to the encoder width
    blocks = [BlockW(b, sh, mode, False) for b in mod.predictor_blocks]
    norm = NormW(mod.predictor_norm, False)
    st.ln_out = torch.empty((B * N, Dp), dtype=mode.torch_dtype, device=dev)
    run.forward(blocks, norm, st.ln_out.data_ptr(), cd)
    st.tgt_rows = torch.empty((B * Kt, Dp), dtype=mode.torch_dtype, device=dev)
    engine.copy_rows(st.ln_out.data_ptr(), cd, Dp, rowmap(Ktv, N, Kcv), st.tgt_rows.data_ptr(), cd, Dp, rowmap(Ktv, Kt, 0),
                     B * Ktv, Dp)
    if Kta > 0:
        engine.copy_rows(st.ln_out.data_ptr(), cd, Dp, rowmap(Kta, N, Kcv + Ktv + Kca), st.tgt_rows.data_ptr(), cd, Dp,
                         rowmap(Kta, Kt, Ktv), B * Kta, Dp)
    proj = mod.predictor_proj
    out = torch.empty((B, Kt, De), dtype=torch.float32, device=dev)
    engine.gemm(mode, GEMM_NT, st.tgt_rows.data_ptr(), sh.weight_ptr(proj.weight, mode), out.data_ptr(), B * Kt, De, Dp,
                Dp, Dp, De, F32, bias=_ptr(proj.bias))
    st.ln_out = None
    if not save:
        st.zc_v = st.zc_a = st.tgt_rows = None
    return out, st


def predictor_backward(mod, parts, st, dout):
    embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a = parts
    run, mode, B, N, mi = st.run, st.mode, st.B, st.N, st.mi
    Kcv, Ktv, Kca, Kta = st.K
    Kt = Ktv + Kta
    De, Dp = embed_v.in_features, embed_v.out_features
    cd, s = mode.code, mode.size
    dev = dout.device
    dout = _f32c(dout)
    lib = _cabi.load()
    al = engine._align
    kmax = max(Kcv, Kca, 1)
    extra = (al(B * Kt * De * s) + al(B * Kt * Dp * s) + al(B * N * Dp * s) + al(B * kmax * Dp * s)
             + 4 * lib.avj_colsum_ws_floats(B * max(Kt, kmax), max(De, Dp)) + (1 << 16))
    sc = engine.SCRATCH.get(run.scratch_bytes() + extra, dev)
    sh = _shadows(mod)
    proj = mod.predictor_proj
    ws = sc.alloc(4 * lib.avj_colsum_ws_floats(B * max(Kt, kmax), max(De, Dp)))
    # ---- predictor_proj backward on the target rows
    dout_c = sc.alloc(B * Kt * De * s)
    engine.copy_rows(dout.data_ptr(), F32, De, IDENTITY, dout_c, cd, De, IDENTITY, B * Kt, De)
    gpw, gpb = engine.grad_ptr(proj.weight), engine.grad_ptr(proj.bias)
    if gpb is not None:
        engine.colsum(dout.data_ptr(), F32, De, IDENTITY, gpb, B * Kt, De, ws)
    if gpw is not None:
        engine.gemm(mode, GEMM_TN, dout_c, st.tgt_rows.data_ptr(), gpw, De, Dp, B * Kt, De, Dp, Dp, F32, accumulate=1)
    d_tgt = sc.alloc(B * Kt * Dp * s)
    engine.gemm(mode, GEMM_NN, dout_c, sh.weight_ptr(proj.weight, mode), d_tgt, B * Kt, Dp, De, De, Dp, Dp, cd)
    # ---- scatter into the gradient of the predictor_norm output (zero on context rows)
    d_ln = sc.alloc(B * N * Dp * s)
    engine.memset0(d_ln, B * N * Dp * s)
    engine.copy_rows(d_tgt, cd, Dp, rowmap(Ktv, Kt, 0), d_ln, cd, Dp, rowmap(Ktv, N, Kcv), B * Ktv, Dp)
    if Kta > 0:
        engine.copy_rows(d_tgt, cd, Dp, rowmap(Kta, Kt, Ktv), d_ln, cd, Dp, rowmap(Kta, N, Kcv + Ktv + Kca), B * Kta, Dp)
    blocks = [BlockW(b, sh, mode, True) for b in mod.predictor_blocks]
    norm = NormW(mod.predictor_norm, True)
    dx0 = run.backward(blocks, norm, d_ln, cd, sc)
    # ---- mask-token gradients: column sums over the target rows
    gt = engine.grad_ptr(tokens_v[mi])
    if gt is not None:
        engine.colsum(dx0, F32, Dp, rowmap(Ktv, N, Kcv), gt, B * Ktv, Dp, ws)
    if Kta > 0:
        gt = engine.grad_ptr(tokens_a[mi])
        if gt is not None:
            engine.colsum(dx0, F32, Dp, rowmap(Kta, N, Kcv + Ktv + Kca), gt, B * Kta, Dp, ws)
    # ---- context embeddings
    dctx = sc.alloc(B * kmax * Dp * s)
    grads_z = []
    for (K, off, zc, emb) in ((Kcv, 0, st.zc_v, embed_v), (Kca, Kcv + Ktv, st.zc_a, embed_a)):
        if K == 0 or emb is None or zc is None:
            grads_z.append(None)
            continue
        rm = rowmap(K, N, off)
        gw, gb = engine.grad_ptr(emb.weight), engine.grad_ptr(emb.bias)
        if gb is not None:
            engine.colsum(dx0, F32, Dp, rm, gb, B * K, Dp, ws)
        engine.copy_rows(dx0, F32, Dp, rm, dctx, cd, Dp, IDENTITY, B * K, Dp)
        if gw is not None:
            engine.gemm(mode, GEMM_TN, dctx, zc.data_ptr(), gw, Dp, De, B * K, Dp, De, De, F32, accumulate=1)
        dz = torch.empty((B, K, De), dtype=torch.float32, device=dev)
        engine.gemm(mode, GEMM_NN, dctx, sh.weight_ptr(emb.weight, mode), dz.data_ptr(), B * K, De, Dp, Dp, De, De, F32)
        grads_z.append(dz)
    return grads_z


class PredictorFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, mod, parts, save, mode, mask_index, z_v, z_a, mcv, mca, mtv, mta, *params):
        out, st = predictor_forward(mod, parts, mask_index, z_v, z_a, mcv, mca, mtv, mta, save, mode)
        ctx.mod, ctx.parts, ctx.st, ctx.n_params = mod, parts, st, len(params)
        ctx.za_shape = None if z_a is None else tuple(z_a.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        dz_v = dz_a = None
        if dout is not None:
            dz_v, dz_a = predictor_backward(ctx.mod, ctx.parts, ctx.st, dout)
            from avjepa_b200 import dist as avj_dist
            sync = avj_dist.active_sync()
            if sync is not None:
                sync.on_backward_done('predictor', ctx.mod)
            if dz_a is None and ctx.za_shape is not None and ctx.needs_input_grad[6]:
                dz_a = torch.zeros(ctx.za_shape, dtype=torch.float32, device=dout.device)
        ctx.st = None
        return (None, None, None, None, None, dz_v, dz_a, None, None, None, None) + (None,) * ctx.n_params


def run_predictor(mod, parts, mask_index, z_v, z_a, mcv, mca, mtv, mta):
    params = [p for p in mod.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and (len(params) > 0 or z_v.requires_grad)
    return PredictorFn.apply(mod, parts, save, engine.Mode.current(), mask_index, z_v, z_a, mcv, mca, mtv, mta, *params)
