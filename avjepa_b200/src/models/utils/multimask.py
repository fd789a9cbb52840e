"""Multi-mask wrappers.

Mirror of ``src/models/utils/multimask.py`` (``MultiMaskWrapper :14-29``, ``AudioVideoMultiMaskWrapper :31-46``,
``PredictorMultiMaskWrapper :49-71``): same constructor, ``.backbone`` attribute (read by ``init_audio_video_model``),
``backbone.`` state-dict prefix, same list-in / list-out forward contract.  The reference loops over the masks and
runs the whole backbone once per mask; here all masks of a call travel through ONE variable-length kernel schedule
(:func:`avjepa_b200.backbone.run_encoder_multi` / ``run_predictor_multi``): every Linear / LayerNorm is a single
launch over the rows of all masks, attention runs per mask.  The results are the same tensors the per-mask loop
produces (``AVJ_MERGE_MASKS=0`` restores the loop for A/B runs).
"""
import torch.nn as nn

from avjepa_b200 import backbone as _bb


def _as_list(v):
    return v if isinstance(v, list) else [v]


def _merged(backbone_module, n_masks):
    return n_masks > 1 and _bb.merge_masks_enabled() and getattr(backbone_module, 'out_layers', None) is None


class MultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, x, masks=None):
        if masks is None:
            return self.backbone(x)
        masks = _as_list(masks)
        if _merged(self.backbone, len(masks)):
            return _bb.run_encoder_multi(self.backbone, x, None, [(m, None) for m in masks])
        return [self.backbone(x, masks=m) for m in masks]


class AudioVideoMultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, x, y, masks=None):
        if masks is None:
            return self.backbone(x, y)
        masks = _as_list(masks)
        if _merged(self.backbone, len(masks)):
            return _bb.run_encoder_multi(self.backbone, x, y, [(m[0], m[1]) for m in masks])
        return [self.backbone(x, y, masks=m) for m in masks]


class PredictorMultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, ctxt, tgt, masks_ctxt, masks_tgt):
        ctxt, tgt, masks_ctxt, masks_tgt = (v if type(v) is list else [v] for v in (ctxt, tgt, masks_ctxt, masks_tgt))
        n = min(len(ctxt), len(tgt), len(masks_ctxt), len(masks_tgt))
        if _merged(self.backbone, n):
            bb = self.backbone
            assert all(mc is not None and mt is not None for mc, mt in zip(masks_ctxt, masks_tgt)), \
                'Cannot run predictor without mask indices'
            if isinstance(ctxt[0], (tuple, list)):                     # audio-video predictor: (video, audio) pairs
                calls = [(i, z[0], z[1], mc[0], mc[1], mt[0], mt[1])
                         for i, (z, mc, mt) in enumerate(zip(ctxt, masks_ctxt, masks_tgt))]
            else:                                                      # video-only predictor
                calls = [(i, z, None, mc, None, mt, None) for i, (z, mc, mt) in enumerate(zip(ctxt, masks_ctxt, masks_tgt))]
            return _bb.run_predictor_multi(bb, bb._parts(), calls)
        return [self.backbone(zi, hi, mc, mt, mask_index=i)
                for i, (zi, hi, mc, mt) in enumerate(zip(ctxt, tgt, masks_ctxt, masks_tgt))]
