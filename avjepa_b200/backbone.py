"""Forward/backward schedules of the encoder and predictor backbones and their autograd nodes.

One ``torch.autograd.Function`` per backbone call: the forward runs patch/ctx embedding and the
whole transformer stack as raw kernel launches and keeps the activations in a
:class:`~avjepa_b200.engine.StackRun`; the backward replays the stack in reverse, ACCUMULATING
parameter gradients straight into ``param.grad`` (created on demand) and returning only the
gradients of tensor inputs.  Reference call sites restated here:

* encoder   -- ``src/models/audiovision_transformer.py:186-239``, ``vision_transformer.py:162-201``
* predictor -- ``src/models/audiovisionpredictor.py:202-301``, ``predictor.py:175-239``
"""
import os

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


def _patchify(x5, idx, B, K, tub, p, mode):
    """[B*K, C*tub*p*p] patch matrix of the kept tokens in the compute dtype (avj_patchify)."""
    _, Cin, T, H, W = x5.shape
    out = torch.empty((B * K, Cin * tub * p * p), dtype=mode.torch_dtype, device=x5.device)
    _cabi.call('avj_patchify', x5.data_ptr(), _ptr(idx), out.data_ptr(), mode.code, B, Cin, T, H, W, tub, p, K, stream())
    return out


def encoder_forward(mod, x, y, masks, save, mode):
    """`masks`: list of (video_idx, audio_idx) pairs, one per mask (entries None = keep every token); `y` and the
    audio indices are None for the video-only encoder.  ALL masks run through one variable-length stack
    (engine.StackRun): sequence group g holds the B sequences of mask g.  Returns ([out_g [B, N_g, D] fp32], state)."""
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
    n_full_a = 0
    if y is not None:
        y = _f32c(y)
        n_full_a = (y.shape[2] // p) * (y.shape[3] // p)
    groups = []
    for mv, ma in masks:
        mv, ma = _idx(mv, dev), _idx(ma, dev)
        Kv = mv.shape[1] if mv is not None else n_full_v
        Ka = 0
        if y is not None:
            Ka = ma.shape[1] if ma is not None else n_full_a
        groups.append((mv, ma, Kv, Ka))
    cd = mode.code
    run = StackRun([(B, Kv + Ka) for _, _, Kv, Ka in groups], D, mod.num_heads, len(mod.blocks), mode, save, dev)
    st = _EncState()
    st.run, st.B, st.mode, st.groups = run, B, mode, groups
    st.N = groups[0][2] + groups[0][3]

    pos_name = 'video_pos_embed' if hasattr(mod, 'video_pos_embed') else 'pos_embed'
    pos_v = _pos_table(mod, x, pos_name)
    pos_a = _f32c(mod.audio_pos_embed) if y is not None else None
    kd_v = Cin * tub * p * p
    pe = mod.patch_embed
    lib = _cabi.load()
    # bf16 mode: im2col-free tf32 kernel (the patch rows go from the clip to the tensor cores by TMA); fp32 check mode
    # and odd geometries: patch matrix + GEMM
    tma_ok = mode.code == _cabi.BF16 and engine.patch_embed_tma_enabled()
    st.patches = []
    for g, (mv, ma, Kv, Ka) in enumerate(groups):
        N = Kv + Ka
        x0 = run.x0_rows(g)
        pv = None
        if tma_ok and lib.avj_patch_embed_supported(p, H, W, T, tub, D):
            engine.patch_embed(x5.data_ptr(), _ptr(mv), _f32c(pe.proj.weight).data_ptr(), x0, B, Cin, T, H, W, tub, p, Kv, D, D,
                               bias=_ptr(pe.proj.bias), pos=pos_v.data_ptr(), pos_idx=_ptr(mv), pos_rows=pos_v.shape[1],
                               out_map=rowmap(Kv, N, 0))
        else:
            pv = _patchify(x5, mv, B, Kv, tub, p, mode)
            engine.gemm(mode, GEMM_NT, pv.data_ptr(), _wptr(pe.proj.weight, mod, mode), x0,
                        B * Kv, D, kd_v, kd_v, kd_v, D, F32, bias=_ptr(pe.proj.bias), pos=pos_v.data_ptr(),
                        pos_idx=_ptr(mv), pos_rows=pos_v.shape[1], out_map=rowmap(Kv, N, 0))
        pa = None
        if y is not None and Ka > 0:
            kd_a = y.shape[1] * p * p
            if tma_ok and lib.avj_patch_embed_supported(p, y.shape[2], y.shape[3], 1, 1, D):
                engine.patch_embed(y.data_ptr(), _ptr(ma), _f32c(pe.audio_proj.weight).data_ptr(), x0, B, y.shape[1], 1, y.shape[2],
                                   y.shape[3], 1, p, Ka, D, D, bias=_ptr(pe.audio_proj.bias), pos=pos_a.data_ptr(),
                                   pos_idx=_ptr(ma), pos_rows=pos_a.shape[1], out_map=rowmap(Ka, N, Kv))
            else:
                pa = _patchify(y.unsqueeze(2), ma, B, Ka, 1, p, mode)
                engine.gemm(mode, GEMM_NT, pa.data_ptr(), _wptr(pe.audio_proj.weight, mod, mode), x0,
                            B * Ka, D, kd_a, kd_a, kd_a, D, F32, bias=_ptr(pe.audio_proj.bias), pos=pos_a.data_ptr(),
                            pos_idx=_ptr(ma), pos_rows=pos_a.shape[1], out_map=rowmap(Ka, N, Kv))
        st.patches.append((pv, pa))
    st.geom = (tub, p)
    st.keep = (x5, y, [g_[:2] for g_ in groups], pos_v, pos_a)       # keep inputs alive until the kernels reading them retire
    sh = _shadows(mod)
    blocks = [BlockW(b, sh, mode, False) for b in mod.blocks]

    if mod.out_layers is not None:
        if save:
            raise NotImplementedError('out_layers is a frozen-eval feature; call under torch.no_grad()')
        if len(groups) != 1:
            raise NotImplementedError('out_layers with several masks in one call')
        N = st.N
        outs = []
        norm = NormW(mod.norm, False)
        for i in range(len(blocks)):
            run.forward_layer(i, blocks[i])
            if i in mod.out_layers:
                o = torch.empty((B, N, D), dtype=torch.float32, device=dev)
                engine.layernorm_fwd(run.x_in(i + 1), norm.w, norm.b, o.data_ptr(), F32, None, None, B * N, D, norm.eps)
                outs.append(o)
        st.run_keepalive = run
        st.out_layers = True
        return outs, st

    outs = [torch.empty((B, Kv + Ka, D), dtype=torch.float32, device=dev) for _, _, Kv, Ka in groups]
    norm = NormW(mod.norm, False) if mod.norm is not None else None
    run.forward(blocks, norm, [o.data_ptr() for o in outs], F32)
    return outs, st


def encoder_backward(mod, st, douts):
    """douts: one gradient per sequence group (None -> zero)."""
    run, mode, B, D = st.run, st.mode, st.B, st.run.D
    cd, s = mode.code, mode.size
    dev = next(d for d in douts if d is not None).device
    douts = [_f32c(d) if d is not None else torch.zeros((B, Kv + Ka, D), dtype=torch.float32, device=dev)
             for d, (_, _, Kv, Ka) in zip(douts, st.groups)]
    lib = _cabi.load()
    kmax = max(max(Kv, Ka, 1) for _, _, Kv, Ka in st.groups)
    extra = engine._align(B * kmax * D * s) * 2 + 4 * lib.avj_colsum_ws_floats(B * kmax, D) + (1 << 16)
    sc = engine.SCRATCH.get(run.scratch_bytes() + extra, dev)
    sh = _shadows(mod)
    blocks = [BlockW(b, sh, mode, True) for b in mod.blocks]
    norm = NormW(mod.norm, True) if mod.norm is not None else None
    from avjepa_b200 import dist as avj_dist
    sync = avj_dist.active_sync()
    evs = sync.layer_events_for(mod, run.L) if sync is not None else None
    dx0 = run.backward(blocks, norm, [d.data_ptr() for d in douts], F32, sc, layer_events=evs)
    if evs is not None:
        sync.on_layers_enqueued(mod, evs)
    pe = mod.patch_embed
    ws = sc.alloc(4 * lib.avj_colsum_ws_floats(B * kmax, D))
    dxc = sc.alloc(B * kmax * D * s)
    for g, (mv, ma, Kv, Ka) in enumerate(st.groups):
        N = Kv + Ka
        dx0_g = dx0 + run.row0[g] * D * 4
        x5, y = st.keep[0], st.keep[1]
        tub, p = st.geom
        for (K, off, patches, conv, src, idx, tb) in ((Kv, 0, st.patches[g][0], pe.proj, x5, mv, tub),
                                                      (Ka, Kv, st.patches[g][1], getattr(pe, 'audio_proj', None),
                                                       y.unsqueeze(2) if y is not None else None, ma, 1)):
            if K == 0 or conv is None or src is None:
                continue
            gw, gb = engine.grad_ptr(conv.weight), engine.grad_ptr(conv.bias) if conv.bias is not None else None
            kd = conv.weight[0].numel()
            fused_wgrad = (patches is None and gw is not None and mode.code == _cabi.BF16
                           and os.environ.get('AVJ_PATCH_WGRAD_GATHER', '1') != '0')     # 0: patch matrix of the kept tokens + GEMM
            if patches is None and gw is not None and not fused_wgrad:
                patches = _patchify(src, idx, B, K, tb, p, mode)
            rm = rowmap(K, N, off)
            if gb is not None:
                engine.colsum(dx0_g, F32, D, rm, gb, B * K, D, ws)
            if gw is not None:
                engine.copy_rows(dx0_g, F32, D, rm, dxc, cd, D, IDENTITY, B * K, D)
                if fused_wgrad:
                    # the forward embedded the tokens straight out of the clip; so does the weight gradient: its second operand
                    # is gathered (and narrowed to bf16) by the GEMM's producer warps, no patch matrix exists in either direction
                    _, Cs, Ts, Hs, Ws = src.shape
                    engine.patch_embed_wgrad(src.data_ptr(), _ptr(idx), dxc, gw, B, Cs, Ts, Hs, Ws, tb, p, K, D)
                else:
                    engine.gemm(mode, GEMM_TN, dxc, patches.data_ptr(), gw, D, kd, B * K, D, kd, kd, F32, accumulate=1)
    if sync is not None:
        sync.on_backward_done('encoder', mod)


class EncoderFn(torch.autograd.Function):
    """autograd node for one encoder call over n_masks masks.  Inputs after the masks are the module's parameters:
    they are listed so autograd knows the outputs depend on them; their gradients are accumulated in place by the
    backward kernels and `None` is returned for them."""

    @staticmethod
    def forward(ctx, mod, save, mode, x, y, n_masks, *rest):
        masks = [(rest[2 * i], rest[2 * i + 1]) for i in range(n_masks)]
        outs, st = encoder_forward(mod, x, y, masks, save, mode)
        ctx.mod, ctx.st, ctx.n_rest = mod, st, len(rest)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        if any(d is not None for d in douts):
            if getattr(ctx.st, 'out_layers', False):
                raise NotImplementedError('out_layers outputs are not differentiable')
            encoder_backward(ctx.mod, ctx.st, list(douts))
        ctx.st = None
        return (None,) * (6 + ctx.n_rest)


def run_encoder_multi(mod, x, y, masks):
    """masks: list of (video_idx, audio_idx) pairs -> list of outputs, all masks in ONE stack schedule."""
    params = [p for p in mod.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and len(params) > 0
    flat = [m for pair in masks for m in pair]
    out = EncoderFn.apply(mod, save, engine.Mode.current(), x, y, len(masks), *flat, *params)
    return list(out)


def run_encoder(mod, x, y, masks_v, masks_a):
    out = run_encoder_multi(mod, x, y, [(masks_v, masks_a)])
    if mod.out_layers is not None:
        return out
    return out[0]


def merge_masks_enabled():
    """AVJ_MERGE_MASKS=0 restores the reference's schedule (one backbone pass per mask) for A/B measurements."""
    import os
    return os.environ.get('AVJ_MERGE_MASKS', '1') != '0'


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
        run = StackRun([(B, N)], D, blocks_mod[0].attn.num_heads, len(blocks_mod), mode, save, x.device,
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


class _PredGroup(object):
    """One mask of a predictor call: its tensors, index sets and row geometry inside the stack."""
    __slots__ = ('mi', 'z_v', 'z_a', 'mcv', 'mca', 'mtv', 'mta', 'Kcv', 'Ktv', 'Kca', 'Kta', 'N', 'Kt', 'zc_v', 'zc_a', 't0')


def predictor_forward(mod, parts, calls, save, mode):
    """parts = (embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a); the *_a entries are None for the video-only
    predictor.  `calls`: list of (mask_index, z_v, z_a, mcv, mca, mtv, mta), one per mask -- all of them run through ONE
    variable-length stack (sequence group g = mask g).  Token layout inside a group: [ctx_v | tgt_v | ctx_a | tgt_a].
    Returns ([out_g [B, Ktv+Kta, D_enc] fp32], state)."""
    embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a = parts
    if tokens_v is None:
        raise NotImplementedError('predictor without mask tokens (the diffusion branch) is not implemented; '
                                  'every AV-JEPA config sets use_mask_tokens=True')
    engine.require_cuda(calls[0][1], 'predictor context')
    dev = calls[0][1].device
    B = calls[0][1].shape[0]
    De, Dp = embed_v.in_features, embed_v.out_features
    cd, s = mode.code, mode.size
    groups = []
    for (mask_index, z_v, z_a, mcv, mca, mtv, mta) in calls:
        g = _PredGroup()
        g.mcv, g.mtv, g.mca, g.mta = _idx(mcv, dev), _idx(mtv, dev), _idx(mca, dev), _idx(mta, dev)
        g.Kcv, g.Ktv = g.mcv.shape[1], g.mtv.shape[1]
        g.Kca = g.mca.shape[1] if (g.mca is not None and z_a is not None) else 0
        g.Kta = g.mta.shape[1] if g.mta is not None else 0
        if z_v.shape[1] != g.Kcv or (z_a is not None and z_a.shape[1] != g.Kca):
            raise ValueError(f'context token count {z_v.shape[1]} does not match the context mask ({g.Kcv})')
        if z_v.shape[0] != B:
            raise ValueError('all masks of one predictor call must share the batch size')
        g.N = g.Kcv + g.Ktv + g.Kca + g.Kta
        g.Kt = g.Ktv + g.Kta
        g.mi = mask_index % len(tokens_v)
        g.z_v, g.z_a = z_v, z_a
        groups.append(g)
    run = StackRun([(B, g.N) for g in groups], Dp, mod.predictor_blocks[0].attn.num_heads, len(mod.predictor_blocks), mode, save, dev)
    st = _PredState()
    st.run, st.B, st.mode, st.groups = run, B, mode, groups
    sh = _shadows(mod)
    pv = _f32c(pos_v)
    pa = _f32c(pos_a) if pos_a is not None else None
    keep = [pv, pa]
    t0 = 0
    for gi, g in enumerate(groups):
        x0, N = run.x0_rows(gi), g.N
        # ---- context rows: x = embed(z) + pos[idx]  (bias and gathered pos fused in the epilogue)
        zk, _, _, zmap_v = _rows_view(g.z_v)
        g.zc_v = torch.empty((B * g.Kcv, De), dtype=mode.torch_dtype, device=dev)
        engine.copy_rows(zk.data_ptr(), F32, De, zmap_v, g.zc_v.data_ptr(), cd, De, IDENTITY, B * g.Kcv, De)
        engine.gemm(mode, GEMM_NT, g.zc_v.data_ptr(), sh.weight_ptr(embed_v.weight, mode), x0, B * g.Kcv, Dp, De, De, De, Dp, F32,
                    bias=_ptr(embed_v.bias), pos=pv.data_ptr(), pos_idx=g.mcv.data_ptr(), pos_rows=pv.shape[1],
                    out_map=rowmap(g.Kcv, N, 0))
        keep.append(zk)
        g.zc_a = None
        if g.Kca > 0:
            zka, _, _, zmap_a = _rows_view(g.z_a)
            g.zc_a = torch.empty((B * g.Kca, De), dtype=mode.torch_dtype, device=dev)
            engine.copy_rows(zka.data_ptr(), F32, De, zmap_a, g.zc_a.data_ptr(), cd, De, IDENTITY, B * g.Kca, De)
            engine.gemm(mode, GEMM_NT, g.zc_a.data_ptr(), sh.weight_ptr(embed_a.weight, mode), x0, B * g.Kca, Dp, De, De, De, Dp,
                        F32, bias=_ptr(embed_a.bias), pos=pa.data_ptr(), pos_idx=g.mca.data_ptr(), pos_rows=pa.shape[1],
                        out_map=rowmap(g.Kca, N, g.Kcv + g.Ktv))
            keep.append(zka)
        # ---- target rows: mask token + pos[idx]
        _cabi.call('avj_fill_mask_tokens', tokens_v[g.mi].data_ptr(), pv.data_ptr(), g.mtv.data_ptr(), x0, Dp,
                   rowmap(g.Ktv, N, g.Kcv), B * g.Ktv, Dp, stream())
        if g.Kta > 0:
            _cabi.call('avj_fill_mask_tokens', tokens_a[g.mi].data_ptr(), pa.data_ptr(), g.mta.data_ptr(), x0, Dp,
                       rowmap(g.Kta, N, g.Kcv + g.Ktv + g.Kca), B * g.Kta, Dp, stream())
        g.z_v = g.z_a = None
        g.t0 = t0                                   # first row of this group inside the packed target-row matrix
        t0 += B * g.Kt
    st.keep = keep
    st.T = t0
    # ---- blocks + predictor_norm over all groups at once, then project the target rows back to the encoder width
    blocks = [BlockW(b, sh, mode, False) for b in mod.predictor_blocks]
    norm = NormW(mod.predictor_norm, False)
    ln_out = torch.empty((run.R, Dp), dtype=mode.torch_dtype, device=dev)
    run.forward(blocks, norm, ln_out.data_ptr(), cd)
    st.tgt_rows = torch.empty((st.T, Dp), dtype=mode.torch_dtype, device=dev)      # [grp0 tgt_v|tgt_a per clip, grp1 ...]
    proj = mod.predictor_proj
    outs = []
    for gi, g in enumerate(groups):
        src = ln_out.data_ptr() + run.row0[gi] * Dp * s
        dst = st.tgt_rows.data_ptr() + g.t0 * Dp * s
        engine.copy_rows(src, cd, Dp, rowmap(g.Ktv, g.N, g.Kcv), dst, cd, Dp, rowmap(g.Ktv, g.Kt, 0), B * g.Ktv, Dp)
        if g.Kta > 0:
            engine.copy_rows(src, cd, Dp, rowmap(g.Kta, g.N, g.Kcv + g.Ktv + g.Kca), dst, cd, Dp, rowmap(g.Kta, g.Kt, g.Ktv),
                             B * g.Kta, Dp)
        out = torch.empty((B, g.Kt, De), dtype=torch.float32, device=dev)
        engine.gemm(mode, GEMM_NT, dst, sh.weight_ptr(proj.weight, mode), out.data_ptr(), B * g.Kt, De, Dp, Dp, Dp, De, F32,
                    bias=_ptr(proj.bias))
        outs.append(out)
    if not save:
        st.tgt_rows = None
        for g in groups:
            g.zc_v = g.zc_a = None
    return outs, st


def predictor_backward(mod, parts, st, douts):
    """douts: one gradient per mask.  Returns [(dz_v, dz_a)] per mask."""
    embed_v, embed_a, tokens_v, tokens_a, pos_v, pos_a = parts
    run, mode, B, groups = st.run, st.mode, st.B, st.groups
    De, Dp = embed_v.in_features, embed_v.out_features
    cd, s = mode.code, mode.size
    dev = next(d for d in douts if d is not None).device
    lib = _cabi.load()
    al = engine._align
    kmax = max(max(g.Kcv, g.Kca, 1) for g in groups)
    T, R = st.T, run.R
    extra = (al(T * De * s) + al(T * Dp * s) + al(R * Dp * s) + al(B * kmax * Dp * s)
             + 4 * lib.avj_colsum_ws_floats(max(T, B * kmax), max(De, Dp)) + (1 << 16))
    sc = engine.SCRATCH.get(run.scratch_bytes() + extra, dev)
    sh = _shadows(mod)
    proj = mod.predictor_proj
    ws = sc.alloc(4 * lib.avj_colsum_ws_floats(max(T, B * kmax), max(De, Dp)))
    # ---- predictor_proj backward on the packed target rows of ALL masks: one bias colsum per mask (fp32 inputs live
    # in separate tensors), then ONE wgrad and ONE dgrad GEMM
    dout_c = sc.alloc(T * De * s)
    gpw, gpb = engine.grad_ptr(proj.weight), engine.grad_ptr(proj.bias)
    for g, d in zip(groups, douts):
        rows = B * g.Kt
        dst = dout_c + g.t0 * De * s
        if d is None:
            engine.memset0(dst, rows * De * s)
            continue
        d = _f32c(d)
        engine.copy_rows(d.data_ptr(), F32, De, IDENTITY, dst, cd, De, IDENTITY, rows, De)
        if gpb is not None:
            engine.colsum(d.data_ptr(), F32, De, IDENTITY, gpb, rows, De, ws)
    if gpw is not None:
        engine.gemm(mode, GEMM_TN, dout_c, st.tgt_rows.data_ptr(), gpw, De, Dp, T, De, Dp, Dp, F32, accumulate=1)
    d_tgt = sc.alloc(T * Dp * s)
    engine.gemm(mode, GEMM_NN, dout_c, sh.weight_ptr(proj.weight, mode), d_tgt, T, Dp, De, De, Dp, Dp, cd)
    # ---- scatter into the gradient of the predictor_norm output (zero on context rows)
    d_ln = sc.alloc(R * Dp * s)
    engine.memset0(d_ln, R * Dp * s)
    for gi, g in enumerate(groups):
        src = d_tgt + g.t0 * Dp * s
        dst = d_ln + run.row0[gi] * Dp * s
        engine.copy_rows(src, cd, Dp, rowmap(g.Ktv, g.Kt, 0), dst, cd, Dp, rowmap(g.Ktv, g.N, g.Kcv), B * g.Ktv, Dp)
        if g.Kta > 0:
            engine.copy_rows(src, cd, Dp, rowmap(g.Kta, g.Kt, g.Ktv), dst, cd, Dp, rowmap(g.Kta, g.N, g.Kcv + g.Ktv + g.Kca),
                             B * g.Kta, Dp)
    blocks = [BlockW(b, sh, mode, True) for b in mod.predictor_blocks]
    norm = NormW(mod.predictor_norm, True)
    dx0_all = run.backward(blocks, norm, d_ln, cd, sc)
    dctx = sc.alloc(B * kmax * Dp * s)
    grads = []
    for gi, g in enumerate(groups):
        dx0 = dx0_all + run.row0[gi] * Dp * 4
        N = g.N
        # ---- mask-token gradients: column sums over the target rows
        gt = engine.grad_ptr(tokens_v[g.mi])
        if gt is not None:
            engine.colsum(dx0, F32, Dp, rowmap(g.Ktv, N, g.Kcv), gt, B * g.Ktv, Dp, ws)
        if g.Kta > 0:
            gt = engine.grad_ptr(tokens_a[g.mi])
            if gt is not None:
                engine.colsum(dx0, F32, Dp, rowmap(g.Kta, N, g.Kcv + g.Ktv + g.Kca), gt, B * g.Kta, Dp, ws)
        # ---- context embeddings
        gz = []
        for (K, off, zc, emb) in ((g.Kcv, 0, g.zc_v, embed_v), (g.Kca, g.Kcv + g.Ktv, g.zc_a, embed_a)):
            if K == 0 or emb is None or zc is None:
                gz.append(None)
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
            gz.append(dz)
        grads.append(tuple(gz))
    return grads


class PredictorFn(torch.autograd.Function):
    """autograd node for one predictor call over n masks; tensor inputs are (z_v, z_a) per mask, then the module's
    parameters (listed for autograd's dependency tracking only)."""

    @staticmethod
    def forward(ctx, mod, parts, save, mode, meta, *rest):
        n = len(meta)
        zs = rest[:2 * n]
        calls = [(meta[i][0], zs[2 * i], zs[2 * i + 1]) + tuple(meta[i][1:]) for i in range(n)]
        outs, st = predictor_forward(mod, parts, calls, save, mode)
        ctx.mod, ctx.parts, ctx.st, ctx.n, ctx.n_rest = mod, parts, st, n, len(rest)
        ctx.za_shapes = [None if zs[2 * i + 1] is None else tuple(zs[2 * i + 1].shape) for i in range(n)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        n = ctx.n
        gz = [None] * (2 * n)
        if any(d is not None for d in douts):
            grads = predictor_backward(ctx.mod, ctx.parts, ctx.st, list(douts))
            from avjepa_b200 import dist as avj_dist
            sync = avj_dist.active_sync()
            if sync is not None:
                sync.on_backward_done('predictor', ctx.mod)
            dev = next(d for d in douts if d is not None).device
            for i, (dz_v, dz_a) in enumerate(grads):
                if dz_a is None and ctx.za_shapes[i] is not None and ctx.needs_input_grad[5 + 2 * i + 1]:
                    dz_a = torch.zeros(ctx.za_shapes[i], dtype=torch.float32, device=dev)
                gz[2 * i], gz[2 * i + 1] = dz_v, dz_a
        ctx.st = None
        return (None, None, None, None, None) + tuple(gz) + (None,) * (ctx.n_rest - 2 * n)


def run_predictor_multi(mod, parts, calls):
    """calls: list of (mask_index, z_v, z_a, mcv, mca, mtv, mta) -> list of outputs, all masks in ONE stack schedule."""
    params = [p for p in mod.parameters() if p.requires_grad]
    save = torch.is_grad_enabled() and (len(params) > 0 or any(c[1].requires_grad for c in calls))
    meta = tuple((c[0], c[3], c[4], c[5], c[6]) for c in calls)
    zs = [t for c in calls for t in (c[1], c[2])]
    return list(PredictorFn.apply(mod, parts, save, engine.Mode.current(), meta, *zs, *params))


def run_predictor(mod, parts, mask_index, z_v, z_a, mcv, mca, mtv, mta):
    return run_predictor_multi(mod, parts, [(mask_index, z_v, z_a, mcv, mca, mtv, mta)])[0]
