"""Frozen-encoder evaluation helpers (drop-in for the reference's ``app/avprediction/utils.py``).

``rebuild_tokens`` (``:206-231``) puts the context tokens and the predicted tokens of every mask back at
their positions in the full token sequence (video tokens first, audio tokens offset by the number of video
tokens).  The reference loops over the batch on the host and scatters with advanced indexing; here each
(mask, tensor) pair is ONE launch of the apply_masks scatter kernel (``avj_gather_rows_bwd``) into a zeroed
fp32 buffer, bit-identical because every destination row receives exactly one source row.
"""
import torch

from avjepa_b200 import _cabi, engine

N_VIDEO_TOKENS = 1568      # 8 x 14 x 14 tubelets of a 16 x 224 x 224 clip (the reference hard-codes the offset)


def rebuild_tokens(ctxt, pred, masks_enc, masks_pred, n_video_tokens=N_VIDEO_TOKENS):
    """ctxt[i] [B, N_ctxt, D], pred[i] [B, N_pred, D], masks_*[i] = (video [B, K_v], audio [B, K_a]) int64.
    Returns a list of fp32 [B, N_ctxt + N_pred, D] tensors (one per mask)."""
    outs = []
    for ctxt_t, pred_t, m_enc, m_pred in zip(ctxt, pred, masks_enc, masks_pred):
        engine.require_cuda(ctxt_t, 'rebuild_tokens')
        B, n_ctxt, D = ctxt_t.shape
        n_pred = pred_t.shape[1]
        n_full = n_ctxt + n_pred
        dev = ctxt_t.device
        full = torch.zeros((B, n_full, D), dtype=torch.float32, device=dev)
        pairs = []
        for src, (m_v, m_a) in ((ctxt_t, m_enc), (pred_t, m_pred)):
            idx = torch.cat((m_v.to(dev), m_a.to(dev) + n_video_tokens), dim=1).to(torch.int64).contiguous()
            if idx.shape[1] != src.shape[1]:
                raise ValueError(f'rebuild_tokens: {idx.shape[1]} indices for {src.shape[1]} tokens')
            pairs.append((src.float().contiguous(), idx))
        hi = max(int(idx.max()) if idx.numel() else -1 for _, idx in pairs)      # eval path: one host sync is fine
        if hi >= n_full:
            raise IndexError(f'index {hi} is out of bounds for dimension 0 with size {n_full}')
        for src, idx in pairs:
            _cabi.call('avj_gather_rows_bwd', _cabi.F32, src.data_ptr(), idx.data_ptr(), full.data_ptr(), B, n_full,
                       idx.shape[1], D, engine.stream())
        outs.append(full)
    return outs
