"""AV-JEPA pre-training step and loop.

Drop-in for the call sites of the reference's ``app/avjepa/train.py``.  The reference keeps its
hot path inside a closure (``train_step``, ``:435-556``) that cannot be imported; here the same
sequence is the importable :class:`TrainStep`:

    schedules -> target fwd (no grad) -> LN + target gather -> 2x context fwd -> 2x predictor
    fwd -> L1 latent loss (+ token-variance reg value) -> backward -> [grad all-reduce] ->
    (clip) -> AdamW -> EMA

with AdamW, gradient unscale/clip, EMA, grad zeroing and the bf16 shadow-weight refresh fused
into one kernel per parameter group, and the two logging-only per-parameter sync loops of the
reference (``grad_logger`` / ``adamw_logger``, ~1000 D2H syncs per step at ViT-L) replaced by
device-side norms.  ``main(args)`` keeps the YAML contract (``:68-159``) and runs the loop on
synthetic clips when ``data.dataset_type == 'synthetic'`` (the video/audio decoders of the
reference's data pipeline are out of scope).
"""
import copy
import logging
import os
import time

import numpy as np
import torch

from avjepa_b200 import dist as avj_dist
from avjepa_b200 import loss as avj_loss
from avjepa_b200.app.avjepa.utils import init_audio_video_model, init_opt, load_checkpoint
from avjepa_b200.backbone import _shadows
from avjepa_b200.optim import FusedAdamWEMA
from avjepa_b200.src.masks.avmultiblock3d import AVMaskCollator as AVMB3DMaskCollator
from avjepa_b200.src.utils.logging import AverageMeter, CSVLogger, get_logger, gpu_timer

_GLOBAL_SEED = 0
log_freq = 10
checkpoint_freq = 1

logger = get_logger(__name__)


def _backbone(m):
    m = getattr(m, 'module', m)
    return getattr(m, 'backbone', m)


class TrainStep(object):
    """One iteration of AV-JEPA pre-training on one GPU (one rank of a data-parallel job)."""

    def __init__(self, encoder, predictor, target_encoder, optimizer, scaler, scheduler, wd_scheduler,
                 momentum_scheduler, loss_exp=1.0, reg_coeff=0.0, clip_grad=None, warmup=40,
                 mixed_precision=True, dtype=torch.bfloat16, grad_sync=None, n_video_tokens=None):
        self.encoder, self.predictor, self.target_encoder = encoder, predictor, target_encoder
        self.optimizer, self.scaler = optimizer, scaler
        self.scheduler, self.wd_scheduler, self.momentum_scheduler = scheduler, wd_scheduler, momentum_scheduler
        self.loss_exp, self.reg_coeff, self.clip_grad, self.warmup = loss_exp, reg_coeff, clip_grad, warmup
        self.mixed_precision, self.dtype = mixed_precision, dtype
        self.grad_sync = grad_sync
        for p in target_encoder.parameters():
            p.requires_grad = False
        if isinstance(optimizer, FusedAdamWEMA):
            optimizer.attach_ema(_backbone(encoder), _backbone(target_encoder))
            for m in (_backbone(encoder), _backbone(predictor), _backbone(target_encoder)):
                optimizer.attach_shadows(m, _shadows(m))
            optimizer.ensure_built()
        self.n_video = n_video_tokens or _backbone(encoder).num_patches
        self.last = {}

    # -- the three forward pieces of the reference closure --------------------------------
    def forward_target(self, clips, asgram, masks_pred_v, masks_pred_a):
        with torch.no_grad():
            h = self.target_encoder(clips, asgram)
            return avj_loss.target_tokens(h, masks_pred_v, masks_pred_a, self.n_video)

    def forward_context(self, clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a):
        masks_enc = list(zip(masks_enc_v, masks_enc_a))
        masks_pred = list(zip(masks_pred_v, masks_pred_a))
        z = self.encoder(clips, asgram, masks_enc)
        z_t = []
        for zi, (mv, ma) in zip(z, masks_enc):
            kv = mv.shape[1]
            z_t.append((zi[:, :kv], zi[:, kv:]))
        return self.predictor(z_t, [None] * len(z_t), masks_enc, masks_pred)

    def __call__(self, clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a, epoch=0, sync=True):
        new_lr = self.scheduler.step()
        new_wd = self.wd_scheduler.step()
        opt = self.optimizer
        with torch.autocast('cuda', dtype=self.dtype, enabled=self.mixed_precision):
            h = self.forward_target(clips, asgram, masks_pred_v, masks_pred_a)
            z = self.forward_context(clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a)
            loss_jepa = avj_loss.jepa_loss(z, h, self.loss_exp, unit_grad=(self.reg_coeff == 0.0))
            if self.reg_coeff != 0.0:
                loss_reg = avj_loss.reg_loss_differentiable(z)
                loss = loss_jepa + self.reg_coeff * loss_reg
            else:
                loss_reg = avj_loss.reg_value(z)
                loss = loss_jepa
        # The loss values exist as soon as the forward has run.  The reference's float(loss) at the end of the
        # step waits for the whole stream (backward, all-reduce, optimizer) and keeps the host from enqueuing the
        # next step; here the three scalars cross to pinned host memory on a side stream NOW and the end of the
        # step only waits for that copy.
        host_vals = self._stage_losses(loss, loss_jepa, loss_reg) if sync else None
        # bf16 needs no loss scaling; the scaler object only mirrors the reference call sites
        if self.grad_sync is not None and hasattr(self.grad_sync, 'begin_step'):
            self.grad_sync.begin_step(opt, _backbone(self.encoder), len(masks_enc_v), len(masks_enc_v))
        loss.backward()
        if isinstance(opt, FusedAdamWEMA):
            opt.mark_grads_dirty()
        # data parallel: ranks SUM their flat gradient buffers; the 1/world averaging is folded into
        # the gradient multiplier the optimizer kernel applies anyway (no extra pass over the grads)
        inv = 1.0
        if self.grad_sync is not None:
            inv = self.grad_sync.all_reduce(opt, average=not isinstance(opt, FusedAdamWEMA))
        enc_norm = pred_norm = None
        coef = None
        if isinstance(opt, FusedAdamWEMA) and (epoch > self.warmup) and (self.clip_grad is not None):
            enc_sq = opt.grad_norm_sq(lambda r: r['group'] in (0, 2))
            pred_sq = opt.grad_norm_sq(lambda r: r['group'] in (1, 3))
            ce, cp = torch.empty_like(enc_sq), torch.empty_like(pred_sq)
            from avjepa_b200 import _cabi, engine
            _cabi.call('avj_clip_coef', enc_sq.data_ptr(), float(self.clip_grad), inv, ce.data_ptr(), engine.stream())
            _cabi.call('avj_clip_coef', pred_sq.data_ptr(), float(self.clip_grad), inv, cp.data_ptr(), engine.stream())
            coef = {0: ce, 2: ce, 1: cp, 3: cp}
            enc_norm, pred_norm = enc_sq.sqrt() * inv, pred_sq.sqrt() * inv
        m = next(self.momentum_scheduler)
        if isinstance(opt, FusedAdamWEMA):
            opt.step(ema_momentum=m, coef_by_group=coef, inv_loss_scale=inv)
            opt.zero_grad()
        else:       # stock optimizer: unfused EMA, reference order
            opt.step()
            opt.zero_grad()
            with torch.no_grad():
                for pq, pk in zip(self.encoder.parameters(), self.target_encoder.parameters()):
                    pk.data.mul_(m).add_((1. - m) * pq.detach().data)
        self.last = dict(loss=loss, loss_jepa=loss_jepa, loss_reg=loss_reg, enc_norm=enc_norm, pred_norm=pred_norm,
                         lr=new_lr, wd=new_wd, momentum=m)
        if not sync:
            return loss, loss_jepa, loss_reg, new_lr, new_wd
        if host_vals is None:
            return float(loss), float(loss_jepa), float(loss_reg), new_lr, new_wd
        buf, ev = host_vals
        ev.synchronize()
        l, lj, lr_ = buf.tolist()
        return l, lj, lr_, new_lr, new_wd

    def _stage_losses(self, loss, loss_jepa, loss_reg):
        """Async D2H of (loss, loss_jepa, loss_reg) behind the forward: returns (pinned buffer, event) or None."""
        if not (torch.is_tensor(loss) and loss.is_cuda):
            return None
        dev = loss.device
        if getattr(self, '_loss_stream', None) is None:
            self._loss_stream = torch.cuda.Stream(device=dev)
            self._loss_host = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_events = [torch.cuda.Event() for _ in range(2)]
            self._loss_slot = 0
        vals = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=dev).detach().reshape(()) for v in (loss, loss_jepa, loss_reg)])
        self._loss_slot ^= 1
        buf, ev = self._loss_host[self._loss_slot], self._loss_events[self._loss_slot]
        self._loss_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._loss_stream):
            buf.copy_(vals, non_blocking=True)
            ev.record(self._loss_stream)
        vals.record_stream(self._loss_stream)
        return buf, ev


def build_training(args, device, world_size=1, rank=0, ipe=None):
    """Everything ``main`` sets up before the loop (reference ``:73-306``), from the YAML dict."""
    meta, mask_cfg, model_cfg = args.get('meta'), args.get('mask'), args.get('model')
    data, opt_cfg, loss_cfg = args.get('data'), args.get('optimization'), args.get('loss')
    which = (meta.get('dtype') or 'float32').lower()
    mixed = which in ('bfloat16', 'float16')
    if which == 'float16':
        raise NotImplementedError('float16 autocast is not implemented; use bfloat16 (every shipped config does)')
    seed = meta.get('seed', _GLOBAL_SEED)
    np.random.seed(seed)
    torch.manual_seed(seed)
    encoder, predictor = init_audio_video_model(
        uniform_power=model_cfg.get('uniform_power', True),
        use_mask_tokens=model_cfg.get('use_mask_tokens', True),
        num_mask_tokens=len(mask_cfg),
        zero_init_mask_tokens=model_cfg.get('zero_init_mask_tokens', True),
        device=device,
        patch_size=data.get('patch_size'),
        num_frames=data.get('num_frames'),
        tubelet_size=data.get('tubelet_size'),
        model_name=model_cfg.get('model_name'),
        crop_size=data.get('crop_size', 224),
        pred_depth=model_cfg.get('pred_depth'),
        pred_embed_dim=model_cfg.get('pred_embed_dim'),
        use_sdpa=meta.get('use_sdpa', False),
    )
    target_encoder = copy.deepcopy(encoder)
    collator = AVMB3DMaskCollator(
        crop_size=data.get('crop_size', 224), num_frames=data.get('num_frames'), patch_size=data.get('patch_size'),
        tubelet_size=data.get('tubelet_size'), cfgs_mask=mask_cfg)
    ipe = ipe or opt_cfg.get('ipe') or 300
    num_epochs, ipe_scale = opt_cfg.get('epochs'), opt_cfg.get('ipe_scale', 1.0)
    optimizer, scaler, scheduler, wd_scheduler = init_opt(
        encoder=encoder, predictor=predictor, wd=float(opt_cfg.get('weight_decay')),
        final_wd=float(opt_cfg.get('final_weight_decay')), start_lr=opt_cfg.get('start_lr'), ref_lr=opt_cfg.get('lr'),
        final_lr=opt_cfg.get('final_lr'), iterations_per_epoch=ipe, warmup=opt_cfg.get('warmup'), num_epochs=num_epochs,
        ipe_scale=ipe_scale, mixed_precision=mixed, betas=opt_cfg.get('betas', (0.9, 0.999)), eps=opt_cfg.get('eps', 1.e-8))
    ema = opt_cfg.get('ema')
    momentum_scheduler = (ema[0] + i * (ema[1] - ema[0]) / (ipe * num_epochs * ipe_scale)
                          for i in range(int(ipe * num_epochs * ipe_scale) + 1))
    grad_sync = avj_dist.GradSync(world_size) if world_size > 1 else None
    step = TrainStep(encoder, predictor, target_encoder, optimizer, scaler, scheduler, wd_scheduler, momentum_scheduler,
                     loss_exp=loss_cfg.get('loss_exp'), reg_coeff=loss_cfg.get('reg_coeff'),
                     clip_grad=opt_cfg.get('clip_grad', None), warmup=opt_cfg.get('warmup'), mixed_precision=mixed,
                     dtype=torch.bfloat16, grad_sync=grad_sync)
    return step, collator, dict(ipe=ipe, num_epochs=num_epochs, batch_size=data.get('batch_size'))


def save_checkpoint(path, step, epoch, loss, batch_size, world_size, lr):
    """Same dict as the reference (``:332-350``), keys included."""
    torch.save({
        'encoder': step.encoder.state_dict(),
        'predictor': step.predictor.state_dict(),
        'opt': step.optimizer.state_dict(),
        'scaler': None if step.scaler is None else step.scaler.state_dict(),
        'target_encoder': step.target_encoder.state_dict(),
        'epoch': epoch, 'loss': loss, 'batch_size': batch_size, 'world_size': world_size, 'lr': lr,
    }, path)


def synthetic_batch(batch_size, generator=None):
    """A collator-shaped batch of synthetic samples: ([clip], label, clip_idx, spectrogram)."""
    return [([torch.randn(3, 16, 224, 224, generator=generator)], 0, [0],
             -80.0 * torch.rand(128, 192, generator=generator)) for _ in range(batch_size)]


def main(args, resume_preempt=False):
    world_size, rank = avj_dist.init_distributed()
    device = torch.device('cuda', torch.cuda.current_device())
    step, collator, info = build_training(args, device, world_size, rank)
    folder, tag = args.get('logging').get('folder'), args.get('logging').get('write_tag')
    os.makedirs(folder, exist_ok=True)
    csv_logger = CSVLogger(os.path.join(folder, f'{tag}_r{rank}.csv'), ('%d', 'epoch'), ('%d', 'itr'), ('%.5f', 'loss'),
                           ('%.5f', 'loss-jepa'), ('%.5f', 'reg-loss'), ('%.5f', 'enc-grad-norm'),
                           ('%.5f', 'pred-grad-norm'), ('%d', 'gpu-time(ms)'), ('%d', 'wall-time(ms)'))
    latest_path = os.path.join(folder, f'{tag}-latest.pth.tar')
    if args.get('data').get('dataset_type', '').lower() != 'synthetic':
        raise NotImplementedError('only data.dataset_type == "synthetic" is available: the decord/ffmpeg/librosa data '
                                  'pipeline of the reference is outside this package (SURVEY.md section 2, row 14)')
    ipe, num_epochs, B = info['ipe'], info['num_epochs'], info['batch_size']
    for epoch in range(num_epochs):
        loss_meter = AverageMeter()
        for itr in range(ipe):
            t0 = time.time()
            while True:
                try:
                    udata, me_v, me_a, mp_v, mp_a = collator(synthetic_batch(B))
                    break
                except TypeError:      # the reference collator's 0-d crash: resample
                    continue
            clips = torch.cat([u.to(device, non_blocking=True) for u in udata[0]], dim=0)
            asgram = udata[3].unsqueeze(1).to(device, non_blocking=True)
            mv = [[m.to(device, non_blocking=True) for m in ms] for ms in (me_v, me_a, mp_v, mp_a)]
            (loss, lj, lr_, new_lr, new_wd), gpu_ms = gpu_timer(lambda: step(clips, asgram, *mv, epoch=epoch))
            loss_meter.update(loss)
            csv_logger.log(epoch + 1, itr, loss, lj, lr_, 0., 0., gpu_ms, (time.time() - t0) * 1000.)
            if itr % log_freq == 0:
                logger.info('[%d, %5d] loss: %.3f [wd: %.2e] [lr: %.2e] [gpu: %.1f ms]' %
                            (epoch + 1, itr, loss_meter.avg, new_wd, new_lr, gpu_ms))
            assert not np.isnan(loss), 'loss is nan'
        if rank == 0 and (epoch % checkpoint_freq == 0 or epoch == num_epochs - 1):
            save_checkpoint(latest_path, step, epoch + 1, loss_meter.avg, B, world_size, new_lr)
