"""Mask application on the device.

Drop-in for the reference's ``src/masks/utils.py``: ``apply_masks(x, masks, concat=True)``
(``:14-34``) gathers the kept token rows of ``x [B, N, D]`` for every ``[B, K]`` index tensor in
``masks`` and concatenates on dim=1 (or returns the list).  The reference expands each index
to a ``[B, K, D]`` int64 tensor and calls ``torch.gather``; here one kernel reads each index
once and moves rows with 16-byte vectors.  Gathered values are bit-identical to the
reference's.  Backward is the matching scatter-add.
"""
import torch

from avjepa_b200 import _cabi, engine


class _GatherRows(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, idx):
        engine.require_cuda(x, 'apply_masks')
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f'apply_masks: unsupported dtype {x.dtype} (float32 / bfloat16)')
        x = x.contiguous()
        idx = idx.to(device=x.device, dtype=torch.int64).contiguous()
        B, N, D = x.shape
        K = idx.shape[1]
        out = torch.empty((B, K, D), dtype=x.dtype, device=x.device)
        code = _cabi.BF16 if x.dtype == torch.bfloat16 else _cabi.F32
        _cabi.call('avj_gather_rows_fwd', code, x.data_ptr(), idx.data_ptr(), out.data_ptr(), B, N, K, D, engine.stream())
        ctx.save_for_backward(idx)
        ctx.shape, ctx.code = (B, N, D), code
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        B, N, D = ctx.shape
        dout = dout.contiguous()
        dx = torch.zeros((B, N, D), dtype=dout.dtype, device=dout.device)
        _cabi.call('avj_gather_rows_bwd', ctx.code, dout.data_ptr(), idx.data_ptr(), dx.data_ptr(), B, N, idx.shape[1], D,
                   engine.stream())
        return dx, None


def apply_masks(x, masks, concat=True):
    """
    :param x: tensor of shape [B (batch-size), N (num-patches), D (feature-dim)]
    :param masks: list of tensors of shape [B, K] containing indices of K patches in [N] to keep
    """
    all_x = [_GatherRows.apply(x, m) for m in masks]
    if not concat:
        return all_x
    return all_x[0] if len(all_x) == 1 else torch.cat(all_x, dim=1)


def get_pred_masks(enc_masks, modal):
    """Complement of each encoder mask in the full token range (1568 video / 96 audio tokens),
    reference ``src/masks/utils.py:49-73``.  Host-side helper; rows keep ascending order."""
    n_full = 1568 if modal == 0 else 96
    out = []
    for m in enc_masks:
        keep = torch.ones((m.shape[0], n_full), dtype=torch.bool, device=m.device)
        keep.scatter_(1, m, False)
        out.append(torch.stack([torch.nonzero(row).flatten() for row in keep]))
    return out
