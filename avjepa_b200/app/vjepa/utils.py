"""Model / optimizer factories of the video-only V-JEPA app.

Drop-in for the reference's ``app/vjepa/utils.py``: ``init_video_model :86-153`` builds the video-only
encoder / predictor pair (``MultiMaskWrapper(VisionTransformer)``, ``PredictorMultiMaskWrapper(
VisionTransformerPredictor)``) with the same arguments, the same second initialisation pass and the same
return tuple; ``init_opt :156-210`` and ``load_checkpoint :28-83`` are line-for-line the functions of the
audio-video app in the reference, so they are shared here (``avjepa_b200.app.avjepa.utils``).
"""
import logging
import sys

import avjepa_b200.src.models.predictor as vit_pred
import avjepa_b200.src.models.vision_transformer as video_vit
from avjepa_b200.app.avjepa.utils import LossScaler, _second_init, init_opt, load_checkpoint  # noqa: F401
from avjepa_b200.src.models.utils.multimask import MultiMaskWrapper, PredictorMultiMaskWrapper

logging.basicConfig(stream=sys.stdout, level=logging.INFO)
logger = logging.getLogger()


def init_video_model(
    device,
    patch_size=16,
    num_frames=16,
    tubelet_size=2,
    model_name='vit_base',
    crop_size=224,
    pred_depth=6,
    pred_embed_dim=384,
    uniform_power=False,
    use_mask_tokens=False,
    num_mask_tokens=2,
    zero_init_mask_tokens=True,
    use_sdpa=False,
):
    encoder = MultiMaskWrapper(video_vit.__dict__[model_name](
        img_size=crop_size, patch_size=patch_size, num_frames=num_frames, tubelet_size=tubelet_size,
        uniform_power=uniform_power, use_sdpa=use_sdpa))
    backbone = encoder.backbone
    predictor = PredictorMultiMaskWrapper(vit_pred.__dict__['vit_predictor'](
        img_size=crop_size, use_mask_tokens=use_mask_tokens, patch_size=patch_size, num_frames=num_frames,
        tubelet_size=tubelet_size, embed_dim=backbone.embed_dim, predictor_embed_dim=pred_embed_dim, depth=pred_depth,
        num_heads=backbone.num_heads, uniform_power=uniform_power, num_mask_tokens=num_mask_tokens,
        zero_init_mask_tokens=zero_init_mask_tokens, use_sdpa=use_sdpa))
    for m in (encoder, predictor):
        _second_init(m)
        m.to(device)
        logger.info(f'{type(m).__name__} number of parameters: {sum(p.numel() for p in m.parameters() if p.requires_grad)}')
    return encoder, predictor
