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
    """Stand-in for ``torch.cuda.amp.GradScaler`` (the reference's ``init_opt`` returns one even for bf16,
    ``app/avjepa/utils.py:281``).  bf16 has fp32's exponent range, so no loss scaling is needed: the scale is the
    constant 1.0, ``scale(loss)`` returns the loss itself and ``unscale_`` has nothing to undo -- which keeps the
    reference's call order ``scale -> backward -> unscale_ -> clip_grad_norm_ -> step -> update``
    (``app/avjepa/train.py:514-523``) exact: clipping between ``unscale_`` and ``step`` sees the true gradients.
    No per-parameter inf-check syncs are issued.  A scale other than 1 (``update(new_scale=...)``) is honoured
    faithfully: ``unscale_`` then divides the optimizer's gradients once, like GradScaler does."""

    def __init__(self, init_scale=1.0, enabled=True):
        self._scale = float(init_scale)
        self._enabled = enabled
        self._unscaled = False

    def is_enabled(self):
        return self._enabled

    def get_scale(self):
        return self._scale if self._enabled else 1.0

    def scale(self, loss):
        if not self._enabled or self._scale == 1.0:
            return loss
        return loss * self._scale

    def unscale_(self, optimizer):
        if self._unscaled:
            raise RuntimeError('unscale_() has already been called on this optimizer since the last update().')
        self._unscaled = True
        if not self._enabled or self._scale == 1.0:
            return
        inv = 1.0 / self._scale
        if isinstance(optimizer, FusedAdamWEMA):
            optimizer.scale_grads(inv)
        else:
            for g in optimizer.param_groups:
                for p in g['params']:
                    if p.grad is not None:
                        p.grad.mul_(inv)

    def step(self, optimizer, *args, **kwargs):
        if not self._unscaled:
            self.unscale_(optimizer)
        return optimizer.step(*args, **kwargs)

    def update(self, new_scale=None):
        self._unscaled = False
        if new_scale is not None:
            self._scale = float(new_scale)

    def state_dict(self):
        return {'scale': self._scale, 'growth_factor': 2.0, 'backoff_factor': 0.5, 'growth_interval': 2000,
                '_growth_tracker': 0}

    def load_state_dict(self, sd):
        """The dynamic scale of a reference checkpoint (2^16 and up) is an artefact of running GradScaler on bf16;
        it carries no training state, so it is NOT adopted -- the scale stays 1."""
        return None


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


def _strip_module(sd):
    return {(k[len('module.'):] if k.startswith('module.') else k): v for k, v in sd.items()}


def _match_keys(module, sd):
    """Reference checkpoints are written from ``nn.DataParallel`` wrappers (``app/avjepa/train.py:298-300,332-350``),
    so their keys are ``module.backbone.*``; this package's modules expose ``backbone.*``.  Re-key `sd` to whatever
    `module` itself uses (works in both directions)."""
    own = list(module.state_dict().keys())
    wants_prefix = bool(own) and all(k.startswith('module.') for k in own)
    sd = _strip_module(sd)
    if wants_prefix:
        sd = {'module.' + k: v for k, v in sd.items()}
    return sd


def load_checkpoint(r_path, encoder, predictor, target_encoder, opt, scaler):
    """Reads a checkpoint written by this package OR by the reference (``app/avjepa/utils.py:28-83``): same
    top-level keys (encoder, predictor, target_encoder, opt, scaler, epoch); the ``module.`` prefix that the
    reference's DataParallel wrappers add is normalised away.  Like the reference, a checkpoint that cannot be
    READ (missing / truncated file) is logged and training starts from epoch 0; a checkpoint whose parameter
    names or shapes do not match the models is an ERROR -- silently restarting from random init would be wrong."""
    epoch = 0
    try:
        checkpoint = torch.load(r_path, map_location=torch.device('cpu'))
    except Exception as e:
        logger.info(f'Encountered exception when loading checkpoint {e}')
        return encoder, predictor, target_encoder, opt, scaler, epoch
    epoch = checkpoint['epoch']
    for name, module in (('encoder', encoder), ('predictor', predictor), ('target_encoder', target_encoder)):
        if module is None:
            continue
        msg = module.load_state_dict(_match_keys(module, checkpoint[name]))      # strict: raises on a mismatch
        logger.info(f'loaded pretrained {name} from epoch {epoch} with msg: {msg}')
    if opt is not None and checkpoint.get('opt') is not None:
        opt.load_state_dict(checkpoint['opt'])
    if scaler is not None and checkpoint.get('scaler') is not None:
        scaler.load_state_dict(checkpoint['scaler'])
    logger.info(f'loaded optimizers from epoch {epoch}')
    logger.info(f'read-path: {r_path}')
    del checkpoint
    return encoder, predictor, target_encoder, opt, scaler, epoch
