"""Fused JEPA latent-prediction loss (forward + backward in one pass over z and h).

Restates ``loss_fn`` / ``reg_fn`` of the reference step (``app/avjepa/train.py:490-498,506-509``):
``loss_jepa = sum_i mean(|z_i - h_i|^p) / p / n_masks`` and
``loss_reg = mean(relu(1 - mean_i sqrt(var_tokens(z_i) + 1e-4)))``.
"""
import torch

from avjepa_b200 import _cabi, engine


class JepaLossFn(torch.autograd.Function):
    """loss = JepaLossFn.apply(loss_exp, mode, beta, unit_grad, z_0..z_{n-1}, h_0..h_{n-1}).

    The kernel that reduces |z-h|^p also writes d loss / d z, so the backward is free; with
    ``unit_grad`` the incoming gradient is taken to be exactly 1 (our own train step calls
    ``loss.backward()`` directly), otherwise dz is rescaled by it."""

    @staticmethod
    def forward(ctx, loss_exp, mode, beta, unit_grad, *zh):
        n = len(zh) // 2
        zs, hs = zh[:n], zh[n:]
        dev = zs[0].device
        engine.require_cuda(zs[0], 'jepa loss')
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        lib = _cabi.load()
        dzs = []
        for z, h in zip(zs, hs):
            z = z.contiguous().float()
            h = h.contiguous().float()
            if z.shape != h.shape:
                raise ValueError(f'jepa loss: prediction {tuple(z.shape)} vs target {tuple(h.shape)}')
            dz = torch.empty_like(z)
            ws = torch.empty(int(lib.avj_loss_ws_floats(z.numel())), dtype=torch.float32, device=dev)
            _cabi.call('avj_loss_fwd_bwd', z.data_ptr(), h.data_ptr(), dz.data_ptr(), out.data_ptr(), z.numel(), n,
                       float(loss_exp), int(mode), float(beta), 1.0, ws.data_ptr(), engine.stream())
            dzs.append(dz)
        ctx.dzs, ctx.unit_grad, ctx.n = dzs, unit_grad, n
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        dzs = ctx.dzs
        if not ctx.unit_grad:
            dzs = [dz * g for dz in dzs]
        return (None, None, None, None) + tuple(dzs) + (None,) * ctx.n


def jepa_loss(z, h, loss_exp=1.0, smooth_l1_beta=None, unit_grad=False):
    """z, h: lists of [B, Kt, D] tensors (one per mask)."""
    mode = 1 if smooth_l1_beta is not None else 0
    return JepaLossFn.apply(float(loss_exp), mode, float(smooth_l1_beta or 0.0), unit_grad, *z, *h)


@torch.no_grad()
def reg_value(z):
    """Value of the token-variance regulariser (logging; reg_coeff is 0 in every shipped config)."""
    B, K, D = z[0].shape
    dev = z[0].device
    pstd = torch.zeros((B, D), dtype=torch.float32, device=dev)
    for zi in z:
        zi = zi.contiguous().float()
        _cabi.call('avj_reg_accumulate', zi.data_ptr(), pstd.data_ptr(), B, zi.shape[1], D, len(z), engine.stream())
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    _cabi.call('avj_reg_finish', pstd.data_ptr(), out.data_ptr(), B * D, engine.stream())
    return out.reshape(())


def reg_loss_differentiable(z):
    """Differentiable form, used only when reg_coeff != 0 (never in the shipped configs)."""
    pstd = sum(torch.sqrt(zi.var(dim=1) + 0.0001) for zi in z) / len(z)
    return torch.mean(torch.nn.functional.relu(1. - pstd))


@torch.no_grad()
def target_tokens(h_full, masks_pred_v, masks_pred_a, n_video):
    """forward_target's tail (``app/avjepa/train.py:448-455``): LayerNorm without affine
    (eps 1e-5) over the target-encoder output, then gather the target rows of every mask and
    lay video|audio targets side by side.  h_full: [B, Nv+Na, D] fp32."""
    B, N, D = h_full.shape
    dev = h_full.device
    h_full = h_full.contiguous()
    hn = torch.empty_like(h_full)
    engine.layernorm_fwd(h_full.data_ptr(), None, None, hn.data_ptr(), _cabi.F32, None, None, B * N, D, 1e-5)
    outs = []
    for mv, ma in zip(masks_pred_v, masks_pred_a if masks_pred_a is not None else [None] * len(masks_pred_v)):
        kv = mv.shape[1]
        ka = ma.shape[1] if ma is not None else 0
        out = torch.empty((B, kv + ka, D), dtype=torch.float32, device=dev)
        _gather_into(hn, mv, 0, out, 0, kv + ka)
        if ka:
            _gather_into(hn, ma, n_video, out, kv, kv + ka)
        outs.append(out)
    return outs


def _gather_into(src, idx, src_row_off, dst, dst_row_off, dst_rows_per_b):
    """dst[b, dst_row_off + j] = src[b, src_row_off + idx[b, j]] -- one gather kernel per segment,
    writing straight into the concatenated destination (no torch.cat)."""
    B, N, D = src.shape
    K = idx.shape[1]
    idx = idx.to(device=src.device, dtype=torch.int64).contiguous()
    if dst_row_off == 0 and dst_rows_per_b == K and src_row_off == 0:
        _cabi.call('avj_gather_rows_fwd', _cabi.F32, src.data_ptr(), idx.data_ptr(), dst.data_ptr(), B, N, K, D, engine.stream())
        return
    # general case: gather to a compact temp then row-mapped copy into place
    tmp = torch.empty((B, K, D), dtype=torch.float32, device=src.device)
    base = src.data_ptr() + src_row_off * D * 4
    _cabi.call('avj_gather_rows_fwd', _cabi.F32, base, idx.data_ptr(), tmp.data_ptr(), B, N, K, D, engine.stream())
    engine.copy_rows(tmp.data_ptr(), _cabi.F32, D, _cabi.IDENTITY, dst.data_ptr(), _cabi.F32, D,
                     engine.rowmap(K, dst_rows_per_b, dst_row_off), B * K, D)
