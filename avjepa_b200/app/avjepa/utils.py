"""Model / optimizer factories of the AV-JEPA app.

Drop-in for the reference's ``app/avjepa/utils.py``: ``init_audio_video_model :86-157``,
``init_opt :228-282``, ``load_checkpoint :28-83`` -- same arguments, same return tuples, same
4 AdamW parameter groups (encoder weights, predictor weights, encoder bias/1-D, predictor
bias/1-D with ``WD_exclude``) and the same second initialisation pass (plain
``trunc_normal_(std=0.02)`` on every Linear AFTER the model's own init, which discards the
depth rescaling -- a reference quirk we reproduce because same-seed parameter parity and
checkpoint interchange depend on it).
"""
import logging
import sys

import torch

import avjepa_b200.src.models.audiovision_transformer as video_vit
import avjepa_b200.src.models.audiovisionpredictor as av_vit_pred
from avjepa_b200.optim import FusedAdamWEMA
from avjepa_b200.src.models.utils.multimask import AudioVideoMultiMaskWrapper, PredictorMultiMaskWrapper
from avjepa_b200.src.utils.schedulers import CosineWDSchedule, WarmupCosineSchedule
from avjepa_b200.src.utils.tensors import trunc_normal_

logging.basicConfig(stream=sys.stdout, level=logging.INFO)
logger = logging.getLogger()


def _second_init(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                torch.nn.init.constant_(m.bias, 0)
        elif isinstance(m, torch.nn.LayerNorm):
            torch.nn.init.constant_(m.bias, 0)
            torch.nn.init.constant_(m.weight, 1.0)


def init_audio_video_model(
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
    encoder = video_vit.__dict__[model_name](
        img_size=crop_size,
        patch_size=patch_size,
        num_frames=num_frames,
        tubelet_size=tubelet_size,
        uniform_power=uniform_power,
        use_sdpa=use_sdpa,
    )
    encoder = AudioVideoMultiMaskWrapper(encoder)
    predictor = av_vit_pred.__dict__['vit_avpredictor'](
        img_size=crop_size,
        use_mask_tokens=use_mask_tokens,
        patch_size=patch_size,
        num_frames=num_frames,
        tubelet_size=tubelet_size,
        embed_dim=encoder.backbone.embed_dim,
        predictor_embed_dim=pred_embed_dim,
        depth=pred_depth,
        num_heads=encoder.backbone.num_heads,
        uniform_power=uniform_power,
        num_mask_tokens=num_mask_tokens,
        zero_init_mask_tokens=zero_init_mask_tokens,
        use_sdpa=use_sdpa,
    )
    predictor = PredictorMultiMaskWrapper(predictor)

    _second_init(encoder)
    _second_init(predictor)

    encoder.to(device)
    predictor.to(device)

    def count_parameters(model):
        return sum(p.numel() for p in model.parameters() if p.requires_grad)

    logger.info(f'Encoder number of parameters: {count_parameters(encoder)}')
    logger.info(f'Predictor number of parameters: {count_parameters(predictor)}')
    return encoder, predictor


class LossScaler(object):
    """Constant-free stand-in for ``torch.cuda.amp.GradScaler`` (reference ``init_opt`` returns one
    even for bf16, ``app/avjepa/utils.py:281``).  bf16 has fp32's exponent range, so scaling by
    2^16 and unscaling is an exact no-op; this object keeps the call sites
    (``scale / unscale_ / step / update / state_dict``) working without the per-parameter
    inf-check syncs, and folds the unscale into the fused AdamW kernel."""

    def __init__(self, init_scale=65536.0, enabled=True):
        self._scale = float(init_scale)
        self._enabled = enabled
        self._unscaled = False

    def is_enabled(self):
        return self._enabled

    def get_scale(self):
        return self._scale

    def scale(self, loss):
        return loss * self._scale if self._enabled else loss

    def unscale_(self, optimizer):
        self._unscaled = True       # folded into the optimizer kernel (see step)

    def step(self, optimizer, *args, **kwargs):
        inv = 1.0 / self._scale if self._enabled else 1.0
        self._unscaled = False
        if isinstance(optimizer, FusedAdamWEMA):
            return optimizer.step(*args, inv_loss_scale=inv, **kwargs)
        if inv != 1.0:
            for g in optimizer.param_groups:
                for p in g['params']:
                    if p.grad is not None:
                        p.grad.mul_(inv)
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None):
        if new_scale is not None:
            self._scale = float(new_scale)

    def state_dict(self):
        return {'scale': self._scale, 'growth_factor': 2.0, 'backoff_factor': 0.5, 'growth_interval': 2000,
                '_growth_tracker': 0}

    def load_state_dict(self, sd):
        self._scale = float(sd.get('scale', self._scale))


def init_opt(
    encoder,
    predictor,
    iterations_per_epoch,
    start_lr,
    ref_lr,
    warmup,
    num_epochs,
    wd=1e-6,
    final_wd=1e-6,
    final_lr=0.0,
    mixed_precision=False,
    ipe_scale=1.25,
    betas=(0.9, 0.999),
    eps=1e-8,
    zero_init_bias_wd=True,
):
    param_groups = [
        {
            'params': (p for n, p in encoder.named_parameters()
                       if ('bias' not in n) and (len(p.shape) != 1))
        }, {
            'params': (p for n, p in predictor.named_parameters()
                       if ('bias' not in n) and (len(p.shape) != 1))
        }, {
            'params': (p for n, p in encoder.named_parameters()
                       if ('bias' in n) or (len(p.shape) == 1)),
            'WD_exclude': zero_init_bias_wd,
            'weight_decay': 0,
        }, {
            'params': (p for n, p in predictor.named_parameters()
                       if ('bias' in n) or (len(p.shape) == 1)),
            'WD_exclude': zero_init_bias_wd,
            'weight_decay': 0,
        },
    ]

    logger.info('Using AdamW (fused AdamW+EMA kernel)')
    optimizer = FusedAdamWEMA(param_groups, betas=betas, eps=eps)
    scheduler = WarmupCosineSchedule(
        optimizer,
        warmup_steps=int(warmup * iterations_per_epoch),
        start_lr=start_lr,
        ref_lr=ref_lr,
        final_lr=final_lr,
        T_max=int(ipe_scale * num_epochs * iterations_per_epoch),
    )
    wd_scheduler = CosineWDSchedule(
        optimizer,
        ref_wd=wd,
        final_wd=final_wd,
        T_max=int(ipe_scale * num_epochs * iterations_per_epoch),
    )
    scaler = LossScaler() if mixed_precision else None
    return optimizer, scaler, scheduler, wd_scheduler


def load_checkpoint(r_path, encoder, predictor, target_encoder, opt, scaler):
    """Reads a checkpoint written by either implementation (same keys: encoder, predictor,
    target_encoder, opt, scaler, epoch).  Errors are logged and swallowed like the reference."""
    epoch = 0
    try:
        checkpoint = torch.load(r_path, map_location=torch.device('cpu'))
        epoch = checkpoint['epoch']
        for name, module in (('encoder', encoder), ('predictor', predictor), ('target_encoder', target_encoder)):
            if module is None:
                continue
            msg = module.load_state_dict(checkpoint[name])
            logger.info(f'loaded pretrained {name} from epoch {epoch} with msg: {msg}')
        opt.load_state_dict(checkpoint['opt'])
        if scaler is not None:
            scaler.load_state_dict(checkpoint['scaler'])
        logger.info(f'loaded optimizers from epoch {epoch}')
        logger.info(f'read-path: {r_path}')
        del checkpoint
    except Exception as e:
        logger.info(f'Encountered exception when loading checkpoint {e}')
        epoch = 0
    return encoder, predictor, target_encoder, opt, scaler, epoch
