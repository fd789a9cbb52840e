"""Multi-mask wrappers: one backbone call per mask.

Mirror of ``src/models/utils/multimask.py`` (``MultiMaskWrapper :14-29``,
``AudioVideoMultiMaskWrapper :31-46``, ``PredictorMultiMaskWrapper :49-71``).  They expose
``.backbone`` (read by ``init_audio_video_model``) and keep the ``backbone.`` state-dict prefix.
"""
import torch.nn as nn


def _as_list(v):
    return v if isinstance(v, list) else [v]


class MultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, x, masks=None):
        if masks is None:
            return self.backbone(x)
        return [self.backbone(x, masks=m) for m in _as_list(masks)]


class AudioVideoMultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, x, y, masks=None):
        if masks is None:
            return self.backbone(x, y)
        return [self.backbone(x, y, masks=m) for m in _as_list(masks)]


class PredictorMultiMaskWrapper(nn.Module):

    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    def forward(self, ctxt, tgt, masks_ctxt, masks_tgt):
        ctxt = ctxt if type(ctxt) is list else [ctxt]
        tgt = tgt if type(tgt) is list else [tgt]
        masks_ctxt = masks_ctxt if type(masks_ctxt) is list else [masks_ctxt]
        masks_tgt = masks_tgt if type(masks_tgt) is list else [masks_tgt]
        return [self.backbone(zi, hi, mc, mt, mask_index=i)
                for i, (zi, hi, mc, mt) in enumerate(zip(ctxt, tgt, masks_ctxt, masks_tgt))]
