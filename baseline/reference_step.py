"""Runs the UNMODIFIED reference (installed into ``baseline/_ref`` by ``tools/install_reference.sh``)
through its own public API and stock PyTorch code path.  MEASUREMENT / TEST INFRASTRUCTURE ONLY -- nothing under
``avjepa_b200/`` imports this file, and no kernel, model or engine of this repository is on this path.

The reference keeps its iteration in a closure (``app/avjepa/train.py:435-556``) that cannot be imported, so
:class:`ReferenceTrainer` restates exactly those lines around the reference's own importable pieces:
``init_audio_video_model`` / ``init_opt`` (``app/avjepa/utils.py:86-157,228-282``), ``apply_masks``
(``src/masks/utils.py``), ``grad_logger`` / ``adamw_logger`` (``src/utils/logging.py:91-118``), wrapped in
``torch.nn.DataParallel`` like ``train.py:298-300`` when a GPU is used.  On the CPU ``torch.cuda.amp.autocast``
and ``GradScaler`` disable themselves (fp32), on a GPU the step runs under bf16 autocast with the GradScaler
the reference creates -- i.e. "stock PyTorch eager on the same box", the competitor SURVEY.md section 8d names.
"""
import copy
import logging
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')


def available():
    return os.path.isdir(os.path.join(REF_DIR, 'src')) and os.path.isdir(os.path.join(REF_DIR, 'app'))


def _import_reference():
    """Import the reference's packages (`src`, `app`) from baseline/_ref."""
    if not available():
        raise RuntimeError(f'{REF_DIR} does not hold the reference; run tools/install_reference.sh in the build container')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    warnings.filterwarnings('ignore', category=FutureWarning)
    lvl = logging.root.manager.disable
    logging.disable(logging.CRITICAL)
    try:
        import app.avjepa.utils as ref_utils
        import src.masks.utils as ref_mask_utils
        import src.utils.logging as ref_logging
        from src.masks.avmultiblock3d import AVMaskCollator
    finally:
        logging.disable(lvl)
    for m in (ref_utils, ref_mask_utils, ref_logging):
        assert os.path.abspath(m.__file__).startswith(REF_DIR), f'{m.__name__} was imported from {m.__file__}, not baseline/_ref'
    return ref_utils, ref_mask_utils, ref_logging, AVMaskCollator


class ReferenceTrainer(object):
    """The reference's models, optimizer, schedulers and restated ``train_step`` on one device."""

    def __init__(self, cfg, device, with_loggers=True):
        import numpy as np
        import torch
        self.torch = torch
        ref_utils, ref_mask_utils, ref_logging, self.AVMaskCollator = _import_reference()
        self.apply_masks = ref_mask_utils.apply_masks
        self.grad_logger, self.adamw_logger = ref_logging.grad_logger, ref_logging.adamw_logger
        meta, mask_cfg, model_cfg = cfg['meta'], cfg['mask'], cfg['model']
        data, opt_cfg, loss_cfg = cfg['data'], cfg['optimization'], cfg['loss']
        self.device = torch.device(device)
        self.on_gpu = self.device.type == 'cuda'
        which = (meta.get('dtype') or 'float32').lower()
        self.mixed_precision = which in ('bfloat16', 'float16')
        self.dtype = torch.bfloat16 if which == 'bfloat16' else (torch.float16 if which == 'float16' else torch.float32)
        seed = meta.get('seed', 0)
        np.random.seed(seed)
        torch.manual_seed(seed)
        lvl = logging.root.manager.disable
        logging.disable(logging.CRITICAL)
        encoder, predictor = ref_utils.init_audio_video_model(
            uniform_power=model_cfg.get('uniform_power', True), use_mask_tokens=model_cfg.get('use_mask_tokens', True),
            num_mask_tokens=len(mask_cfg), zero_init_mask_tokens=model_cfg.get('zero_init_mask_tokens', True),
            device=self.device, patch_size=data['patch_size'], num_frames=data['num_frames'], tubelet_size=data['tubelet_size'],
            model_name=model_cfg['model_name'], crop_size=data.get('crop_size', 224), pred_depth=model_cfg['pred_depth'],
            pred_embed_dim=model_cfg['pred_embed_dim'], use_sdpa=meta.get('use_sdpa', False))
        target_encoder = copy.deepcopy(encoder)
        ipe = opt_cfg.get('ipe') or 300
        self.num_epochs, ipe_scale = opt_cfg['epochs'], opt_cfg.get('ipe_scale', 1.0)
        self.optimizer, self.scaler, self.scheduler, self.wd_scheduler = ref_utils.init_opt(
            encoder=encoder, predictor=predictor, wd=float(opt_cfg['weight_decay']), final_wd=float(opt_cfg['final_weight_decay']),
            start_lr=opt_cfg['start_lr'], ref_lr=opt_cfg['lr'], final_lr=opt_cfg['final_lr'], iterations_per_epoch=ipe,
            warmup=opt_cfg['warmup'], num_epochs=self.num_epochs, ipe_scale=ipe_scale, mixed_precision=self.mixed_precision,
            betas=opt_cfg.get('betas', (0.9, 0.999)), eps=opt_cfg.get('eps', 1.e-8))
        logging.disable(lvl)
        if self.on_gpu:                                           # train.py:298-300
            ids = [self.device.index if self.device.index is not None else torch.cuda.current_device()]
            encoder = torch.nn.DataParallel(encoder, device_ids=ids)
            predictor = torch.nn.DataParallel(predictor, device_ids=ids)
            target_encoder = torch.nn.DataParallel(target_encoder, device_ids=ids)
        for p in target_encoder.parameters():
            p.requires_grad = False
        self.encoder, self.predictor, self.target_encoder = encoder, predictor, target_encoder
        ema = opt_cfg['ema']
        self.momentum_scheduler = (ema[0] + i * (ema[1] - ema[0]) / (ipe * self.num_epochs * ipe_scale)
                                   for i in range(int(ipe * self.num_epochs * ipe_scale) + 1))
        self.loss_exp, self.reg_coeff = loss_cfg['loss_exp'], loss_cfg['reg_coeff']
        self.clip_grad, self.warmup = opt_cfg.get('clip_grad', None), opt_cfg['warmup']
        self.with_loggers = with_loggers
        self.mask_cfg, self.data_cfg = mask_cfg, data

    def collator(self):
        d = self.data_cfg
        return self.AVMaskCollator(crop_size=d.get('crop_size', 224), num_frames=d['num_frames'], patch_size=d['patch_size'],
                                   tubelet_size=d['tubelet_size'], cfgs_mask=self.mask_cfg)

    def train_step(self, clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a, epoch=0, keep_grads=False):
        """``app/avjepa/train.py:435-556`` restated line for line around the reference modules."""
        torch = self.torch
        F = torch.nn.functional
        encoder, predictor, target_encoder = self.encoder, self.predictor, self.target_encoder
        optimizer, scaler = self.optimizer, self.scaler
        _new_lr = self.scheduler.step()
        _new_wd = self.wd_scheduler.step()

        def forward_target(c, a):
            with torch.no_grad():
                h = target_encoder(c, a)
                h = F.layer_norm(h, (h.size(-1),))
                video_tokens, audio_tokens = torch.split(h, [1568, 96], dim=1)
                h_v = self.apply_masks(video_tokens, masks_pred_v, concat=False)
                h_a = self.apply_masks(audio_tokens, masks_pred_a, concat=False)
                out = [torch.cat([h_v[i], h_a[i]], dim=1) for i in range(len(h_v))]
                return h_v, h_a, out

        def forward_context(c, a, h_v, h_a):
            masks_enc = list(zip(masks_enc_v, masks_enc_a))
            masks_pred = list(zip(masks_pred_v, masks_pred_a))
            h = list(zip(h_v, h_a))
            z = encoder(c, a, masks_enc)
            z_t = []
            for zi, (mv, ma) in zip(z, masks_enc):
                z_t.append(torch.split(zi, [mv.shape[1], ma.shape[1]], dim=1))
            return predictor(z_t, h, masks_enc, masks_pred)

        def loss_fn(z, h):
            loss = 0.
            for zi, hi in zip(z, h):
                loss += torch.mean(torch.abs(zi - hi) ** self.loss_exp) / self.loss_exp
            loss /= len(masks_pred_v)
            return loss

        def reg_fn(z):
            return sum([torch.sqrt(zi.var(dim=1) + 0.0001) for zi in z]) / len(z)

        loss_reg = 0.
        with torch.cuda.amp.autocast(dtype=self.dtype, enabled=self.mixed_precision):
            h_v, h_a, h = forward_target(clips, asgram)
            z = forward_context(clips, asgram, h_v, h_a)
            loss_jepa = loss_fn(z, h)
            pstd_z = reg_fn(z)
            loss_reg += torch.mean(F.relu(1. - pstd_z))
        loss = loss_jepa + self.reg_coeff * loss_reg

        _enc_norm, _pred_norm = 0., 0.
        if self.mixed_precision:
            scaler.scale(loss).backward()
            scaler.unscale_(optimizer)
        else:
            loss.backward()
        if (epoch > self.warmup) and (self.clip_grad is not None):
            _enc_norm = torch.nn.utils.clip_grad_norm_(encoder.parameters(), self.clip_grad)
            _pred_norm = torch.nn.utils.clip_grad_norm_(predictor.parameters(), self.clip_grad)
        grads = None
        if keep_grads:
            grads = {}
            for tag, m in (('enc', encoder), ('pred', predictor)):
                for n, p in m.named_parameters():
                    if p.grad is not None:
                        grads[tag + '.' + n.replace('module.', '', 1).replace('backbone.', '', 1)] = p.grad.detach().float().cpu().clone()
        if self.mixed_precision:
            scaler.step(optimizer)
            scaler.update()
        else:
            optimizer.step()
        grad_stats = grad_stats_pred = optim_stats = None
        if self.with_loggers:
            grad_stats = self.grad_logger(encoder.named_parameters())
            grad_stats.global_norm = float(_enc_norm)
            grad_stats_pred = self.grad_logger(predictor.named_parameters())
            grad_stats_pred.global_norm = float(_pred_norm)
        optimizer.zero_grad()
        if self.with_loggers:
            optim_stats = self.adamw_logger(optimizer)

        m = next(self.momentum_scheduler)
        with torch.no_grad():
            for param_q, param_k in zip(encoder.parameters(), target_encoder.parameters()):
                param_k.data.mul_(m).add_((1. - m) * param_q.detach().data)

        out = dict(loss=float(loss.detach()), loss_jepa=float(loss_jepa.detach()), loss_reg=float(loss_reg.detach()), lr=_new_lr, wd=_new_wd, momentum=m,
                   enc_norm=float(_enc_norm), pred_norm=float(_pred_norm), grad_stats=grad_stats,
                   grad_stats_pred=grad_stats_pred, optim_stats=optim_stats)
        if keep_grads:
            out['grads'] = grads
        return out

    def state_for_product(self):
        """{'encoder': state_dict, 'predictor': state_dict} with the DataParallel `module.` prefix removed -- used to
        give the product modules bit-identical weights."""
        strip = lambda sd: {k.replace('module.', '', 1): v.detach().clone() for k, v in sd.items()}  # noqa: E731
        return dict(encoder=strip(self.encoder.state_dict()), predictor=strip(self.predictor.state_dict()))
