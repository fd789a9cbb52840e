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
from avjepa_b200.src.utils.distributed import AllReduce
from avjepa_b200.src.utils.logging import AverageMeter, CSVLogger, DeviceParamStats, get_logger, gpu_timer

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
                 mixed_precision=True, dtype=torch.bfloat16, grad_sync=None, n_video_tokens=None, smooth_l1_beta=None):
        self.encoder, self.predictor, self.target_encoder = encoder, predictor, target_encoder
        self.optimizer, self.scaler = optimizer, scaler
        self.scheduler, self.wd_scheduler, self.momentum_scheduler = scheduler, wd_scheduler, momentum_scheduler
        self.loss_exp, self.reg_coeff, self.clip_grad, self.warmup = loss_exp, reg_coeff, clip_grad, warmup
        self.mixed_precision, self.dtype = mixed_precision, dtype
        self.grad_sync = grad_sync
        self.smooth_l1_beta = smooth_l1_beta      # None: the reference's |z-h|^p / p;  float: smooth-L1 with this beta
        self.param_stats = DeviceParamStats()
        self.keep_zh, self.last_zh = False, None
        self.stats = None
        self.pipeline_optimizer = os.environ.get('AVJ_DDP_PIPELINE_OPT', '1') != '0'
        # AVJ_TARGET_STREAM=1: target encoder on a second stream beside the context / predictor forward.  Measured neutral
        # (213.9 vs 213.9 clips/s, ViT-L B=24: the persistent GEMMs of either stream already fill the machine), so the
        # default keeps the reference's order on one stream.
        self.target_stream_enabled = os.environ.get('AVJ_TARGET_STREAM', '0') == '1'
        self._target_stream = None
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

    def forward_loss(self, clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a):
        """Step 1 of the reference closure (``:500-509``): (loss, loss_jepa, loss_reg) as device scalars."""
        with torch.autocast('cuda', dtype=self.dtype, enabled=self.mixed_precision):
            if self.target_stream_enabled and clips.is_cuda:
                # The stop-gradient target pass is independent of the context / predictor forward until the loss: it runs
                # on a second stream, so its large GEMMs fill the partial waves and memory-bound stretches (LayerNorm,
                # gathers, small-sequence attention) of the context pass and vice versa.
                main = torch.cuda.current_stream()
                if self._target_stream is None:
                    self._target_stream = torch.cuda.Stream(device=clips.device)
                side = self._target_stream
                side.wait_stream(main)                     # inputs and the EMA-updated target weights are ready
                with torch.cuda.stream(side):
                    h = self.forward_target(clips, asgram, masks_pred_v, masks_pred_a)
                z = self.forward_context(clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a)
                main.wait_stream(side)
                for t in h:
                    t.record_stream(main)                  # allocated on the side stream, consumed by the loss here
            else:
                h = self.forward_target(clips, asgram, masks_pred_v, masks_pred_a)
                z = self.forward_context(clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a)
            loss_jepa = avj_loss.jepa_loss(z, h, self.loss_exp, smooth_l1_beta=self.smooth_l1_beta,
                                           unit_grad=(self.reg_coeff == 0.0))
            if self.reg_coeff != 0.0:
                loss_reg = avj_loss.reg_loss_differentiable(z)
                loss = loss_jepa + self.reg_coeff * loss_reg
            else:
                loss_reg = avj_loss.reg_value(z)
                loss = loss_jepa
        if self.keep_zh:               # tests / parity probes only: predictions and targets of this step
            self.last_zh = ([t.detach() for t in z], h)
        return loss, loss_jepa, loss_reg

    def backward_and_reduce(self, loss, n_masks=2, defer=False):
        """Backward, then (data parallel) the SUM all-reduce of the flat gradient buffers -- overlapped with the
        backward when the GradSync supports it.  Returns the factor still owed to the gradients (1/world when the
        fused optimizer folds the averaging into its gradient multiplier, else 1.0).  bf16 needs no loss scaling
        (LossScaler is the identity); the scaler object only mirrors the reference call sites."""
        opt = self.optimizer
        if self.grad_sync is not None and hasattr(self.grad_sync, 'begin_step'):
            # how many backward calls of each backbone this step makes: ONE when the MultiMask wrappers run all
            # masks through a single variable-length schedule (the default), else one per mask
            from avjepa_b200.backbone import merge_masks_enabled
            n_calls = 1 if merge_masks_enabled() else n_masks
            self.grad_sync.begin_step(opt, _backbone(self.encoder), n_calls, n_calls)
        loss.backward()
        if self.grad_sync is None:
            return 1.0
        if defer and hasattr(self.grad_sync, 'can_pipeline') and self.grad_sync.can_pipeline(opt):
            return None                  # the caller finishes with grad_sync.finish_pipelined (optimizer per reduced interval)
        return self.grad_sync.all_reduce(opt, average=not isinstance(opt, FusedAdamWEMA))

    def __call__(self, clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a, epoch=0, sync=True,
                 log_stats=False):
        """One iteration.  Returns ``(loss, loss_jepa, loss_reg, lr, wd)`` like the reference closure's first five
        values (floats when `sync`, device scalars otherwise).  With `log_stats` the per-parameter gradient-norm /
        Adam-moment statistics of ``grad_logger`` / ``adamw_logger`` (``app/avjepa/train.py:526-531``) are computed
        on the device as well and left in ``self.stats`` = (enc grad stats, pred grad stats, optim stats)."""
        new_lr = self.scheduler.step()
        new_wd = self.wd_scheduler.step()
        opt = self.optimizer
        fused = isinstance(opt, FusedAdamWEMA)
        loss, loss_jepa, loss_reg = self.forward_loss(clips, asgram, masks_enc_v, masks_enc_a, masks_pred_v, masks_pred_a)
        # The loss values exist as soon as the forward has run.  The reference's float(loss) at the end of the step
        # waits for the whole stream (backward, all-reduce, optimizer) and keeps the host from enqueuing the next
        # step; here the three scalars cross to pinned host memory on a side stream NOW and the end of the step only
        # waits for that copy (unless gradient norms / parameter statistics, which exist only after backward, were
        # asked for as well).
        early = self._stage(0, [loss, loss_jepa, loss_reg]) if sync else None
        # Without clipping and statistics nothing needs ALL gradients at once: every gradient interval can be updated as
        # soon as its own all-reduce is done, under the all-reduces of the intervals behind it (AVJ_DDP_PIPELINE_OPT=0: off).
        clipping = (epoch > self.warmup) and (self.clip_grad is not None)
        defer = fused and not clipping and not log_stats and self.pipeline_optimizer
        inv = self.backward_and_reduce(loss, n_masks=len(masks_enc_v), defer=defer)
        enc_norm = pred_norm = None
        coef = None
        pipelined = inv is None
        if pipelined:
            m = next(self.momentum_scheduler)
            inv = 1.0 / self.grad_sync.world_size
            opt.begin_step()
            for g, lo, hi in self.grad_sync.finish_pipelined(opt):
                opt.step_interval(opt.range_of_grad(g), lo, hi, ema_momentum=m, inv_loss_scale=inv, zero_grads=True)
            for r in opt._ranges:                                   # ranges without gradients (frozen tables): EMA copy only
                if r['g'] is None:
                    opt.step_interval(r, 0, r['flat'].numel(), ema_momentum=m, inv_loss_scale=inv, zero_grads=True)
            opt.end_step(zero_grads=True)
        elif fused and (epoch > self.warmup) and (self.clip_grad is not None):       # train.py:518-520
            enc_sq = opt.grad_norm_sq(lambda r: r['group'] in (0, 2))
            pred_sq = opt.grad_norm_sq(lambda r: r['group'] in (1, 3))
            ce, cp = torch.empty_like(enc_sq), torch.empty_like(pred_sq)
            from avjepa_b200 import _cabi, engine
            _cabi.call('avj_clip_coef', enc_sq.data_ptr(), float(self.clip_grad), inv, ce.data_ptr(), engine.stream())
            _cabi.call('avj_clip_coef', pred_sq.data_ptr(), float(self.clip_grad), inv, cp.data_ptr(), engine.stream())
            coef = {0: ce, 2: ce, 1: cp, 3: cp}
            enc_norm, pred_norm = enc_sq.sqrt() * inv, pred_sq.sqrt() * inv
        elif (not fused) and (epoch > self.warmup) and (self.clip_grad is not None):
            enc_norm = torch.nn.utils.clip_grad_norm_(self.encoder.parameters(), self.clip_grad)
            pred_norm = torch.nn.utils.clip_grad_norm_(self.predictor.parameters(), self.clip_grad)
        if not pipelined:
            m = next(self.momentum_scheduler)
        if pipelined:
            pass
        elif fused:
            if log_stats:
                self.param_stats.capture_grads(opt, coef, scale=inv)     # before the kernel below zeroes them
            opt.step(ema_momentum=m, coef_by_group=coef, inv_loss_scale=inv, zero_grads=True)
            if log_stats:
                self.param_stats.capture_moments(opt)
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
        buf, ev = early
        ev.synchronize()
        l, lj, lr_ = buf.tolist()
        want_norms = enc_norm is not None
        self.stats = None
        if want_norms or (fused and log_stats):
            # the two global gradient norms and (with log_stats) the per-parameter statistics exist only after the
            # backward: ONE more asynchronous copy carries all of them (the reference blocks on float(norm) x2 and
            # ~1000 per-parameter float() calls here)
            zero = torch.zeros((), dtype=torch.float64, device=loss.device)
            if not (fused and log_stats):
                self.param_stats._g = self.param_stats._m = self.param_stats._v = None
            self.param_stats.stage(self._copy_stream, extra=[enc_norm if want_norms else zero, pred_norm if want_norms else zero])
            enc_stats, pred_stats, optim_stats, (en, pn) = self.param_stats.collect(
                list(self.encoder.named_parameters()), list(self.predictor.named_parameters()))
            self.last.update(enc_norm=en, pred_norm=pn)
            if log_stats:
                enc_stats.global_norm, pred_stats.global_norm = en, pn
                self.stats = (enc_stats, pred_stats, optim_stats)
        return l, lj, lr_, new_lr, new_wd


    def _stage(self, slot, values):
        """Asynchronous D2H of a few device scalars on the copy stream: returns (pinned fp32 buffer, event)."""
        dev = values[0].device if torch.is_tensor(values[0]) else torch.device('cuda', torch.cuda.current_device())
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = {}
        key = (slot, len(values))
        if key not in self._slots:
            self._slots[key] = [[torch.empty(len(values), dtype=torch.float32).pin_memory(), torch.cuda.Event()] for _ in range(2)]
            self._slots[key].append(0)
        ring = self._slots[key]
        ring[2] ^= 1
        buf, ev = ring[ring[2]]
        vals = torch.stack([torch.as_tensor(v, dtype=torch.float32, device=dev).detach().reshape(()) for v in values])
        self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy_stream):
            buf.copy_(vals, non_blocking=True)
            ev.record(self._copy_stream)
        vals.record_stream(self._copy_stream)
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
    if meta.get('device_masks', False) and torch.device(device).type == 'cuda':
        # meta.device_masks: block positions drawn and index sets built on the GPU from a replica of torch's CPU generator
        # (bit-identical masks, no mask H2D copies); one call ahead on a side stream
        from avjepa_b200.src.masks.device_collator import DeviceAVMaskCollator
        collator = DeviceAVMaskCollator(
            crop_size=data.get('crop_size', 224), num_frames=data.get('num_frames'), patch_size=data.get('patch_size'),
            tubelet_size=data.get('tubelet_size'), cfgs_mask=mask_cfg, device=device, prefetch=True)
    else:
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


def save_checkpoint(path, step, epoch, loss, batch_size, world_size, lr, reference_keys=False):
    """Same dict as the reference (``:332-350``), keys included.  `reference_keys`: write the model state dicts with
    the ``module.`` prefix the reference's ``nn.DataParallel`` wrappers produce, so the reference's own (strict)
    ``load_checkpoint`` reads the file; :func:`load_checkpoint` here reads either form."""
    def sd(m):
        d = m.state_dict()
        return {('module.' + k if reference_keys and not k.startswith('module.') else k): v for k, v in d.items()}
    torch.save({
        'encoder': sd(step.encoder),
        'predictor': sd(step.predictor),
        'opt': step.optimizer.state_dict(),
        'scaler': None if step.scaler is None else step.scaler.state_dict(),
        'target_encoder': sd(step.target_encoder),
        'epoch': epoch, 'loss': loss, 'batch_size': batch_size, 'world_size': world_size, 'lr': lr,
    }, path)


def synthetic_batch(batch_size, generator=None):
    """A collator-shaped batch of synthetic samples: ([clip], label, clip_idx, spectrogram)."""
    return [([torch.randn(3, 16, 224, 224, generator=generator)], 0, [0],
             -80.0 * torch.rand(128, 192, generator=generator)) for _ in range(batch_size)]


def main(args, resume_preempt=False):
    """The reference loop (``app/avjepa/train.py:68-644``) around :class:`TrainStep`: YAML dict in, CSV rows and
    ``<tag>-latest.pth.tar`` out, resume from ``meta.load_checkpoint`` / ``meta.read_checkpoint`` or an existing
    latest checkpoint with the schedulers, momentum generator and mask collator fast-forwarded (``:309-330``)."""
    world_size, rank = avj_dist.init_distributed()
    device = torch.device('cuda', torch.cuda.current_device())
    step, collator, info = build_training(args, device, world_size, rank)
    meta = args.get('meta')
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

    # -- resume (reference :187-193, :309-330)
    load_model = bool(meta.get('load_checkpoint', False)) or resume_preempt
    r_file = meta.get('read_checkpoint', None)
    load_path = None
    if load_model:
        load_path = os.path.join(folder, r_file) if r_file is not None else latest_path
        if not os.path.exists(load_path):
            load_path, load_model = None, False
    start_epoch = 0
    if load_model or os.path.exists(latest_path):
        logger.info('LOADING CHECKPOINTS')
        *_, start_epoch = load_checkpoint(r_path=load_path or latest_path, encoder=step.encoder, predictor=step.predictor,
                                          target_encoder=step.target_encoder, opt=step.optimizer, scaler=step.scaler)
        for _ in range(start_epoch * ipe):
            step.scheduler.step()
            step.wd_scheduler.step()
            next(step.momentum_scheduler)
            collator.step()

    # every rank draws its own clips (the reference shards the dataset with a DistributedSampler)
    data_gen = torch.Generator().manual_seed(int(meta.get('seed', _GLOBAL_SEED)) * 1000003 + rank)
    new_lr = new_wd = 0.
    for epoch in range(start_epoch, num_epochs):
        loss_meter, input_var_meter, input_var_min_meter = AverageMeter(), AverageMeter(), AverageMeter()
        jepa_loss_meter, reg_loss_meter = AverageMeter(), AverageMeter()
        gpu_time_meter, wall_time_meter = AverageMeter(), AverageMeter()
        mask_meters = [AverageMeter() for _ in range(len(args.get('mask')))]
        for itr in range(ipe):
            t0 = time.time()
            while True:
                try:
                    udata, me_v, me_a, mp_v, mp_a = collator(synthetic_batch(B, data_gen))
                    break
                except TypeError:      # the reference collator's 0-d crash: resample
                    continue
            clips = torch.cat([u.to(device, non_blocking=True) for u in udata[0]], dim=0)
            asgram = udata[3].unsqueeze(1).to(device, non_blocking=True)
            mv = [[m.to(device, non_blocking=True) for m in ms] for ms in (me_v, me_a, mp_v, mp_a)]
            for i, mm in enumerate(mask_meters):
                mm.update(me_v[i][0].size(-1))
            want_stats = (itr % log_freq == 0)
            (loss, lj, lr_, new_lr, new_wd), gpu_ms = gpu_timer(
                lambda: step(clips, asgram, *mv, epoch=epoch, log_stats=want_stats))
            wall_ms = (time.time() - t0) * 1000.
            loss_meter.update(loss)
            # the two logging collectives of the reference (:560-561)
            per_clip_var = clips.view(clips.shape[0], -1).var(dim=1)
            input_var = float(AllReduce.apply(per_clip_var.mean(dim=0)))
            input_var_min = float(AllReduce.apply(torch.min(per_clip_var)))
            input_var_meter.update(input_var)
            input_var_min_meter.update(input_var_min)
            jepa_loss_meter.update(lj)
            reg_loss_meter.update(lr_)
            gpu_time_meter.update(gpu_ms)
            wall_time_meter.update(wall_ms)
            enc_norm, pred_norm = step.last.get('enc_norm') or 0., step.last.get('pred_norm') or 0.
            csv_logger.log(epoch + 1, itr, loss, lj, lr_, enc_norm, pred_norm, gpu_ms, wall_ms)
            if want_stats or np.isnan(loss) or np.isinf(loss):
                logger.info('[%d, %5d] loss: %.3f | p%.3f r%.3f | input_var: %.3f %.3f | masks: %s [wd: %.2e] [lr: %.2e] '
                            '[mem: %.2e] [gpu: %.1f ms][wall: %.1f ms]'
                            % (epoch + 1, itr, loss_meter.avg, jepa_loss_meter.avg, reg_loss_meter.avg, input_var_meter.avg,
                               input_var_min_meter.avg, '[' + ', '.join('%.1f' % m.avg for m in mask_meters) + ']',
                               new_wd, new_lr, torch.cuda.max_memory_allocated() / 1024.0 ** 2, gpu_time_meter.avg,
                               wall_time_meter.avg))
                if step.stats is not None:
                    g_enc, g_pred, o = step.stats
                    logger.info('[%d, %5d] first moment: %.2e [%.2e %.2e] second moment: %.2e [%.2e %.2e]'
                                % (epoch + 1, itr, o['exp_avg'].avg, o['exp_avg'].min, o['exp_avg'].max,
                                   o['exp_avg_sq'].avg, o['exp_avg_sq'].min, o['exp_avg_sq'].max))
                    for nm, g in (('enc', g_enc), ('pred', g_pred)):
                        logger.info('[%d, %5d] %s_grad_stats: f/l[%.2e %.2e] mn/mx(%.2e, %.2e) %.2e'
                                    % (epoch + 1, itr, nm, g.first_layer, g.last_layer, g.min, g.max, g.global_norm))
            assert not np.isnan(loss), 'loss is nan'
        logger.info('avg. loss %.3f' % loss_meter.avg)
        if rank == 0 and (epoch % checkpoint_freq == 0 or epoch == num_epochs - 1):
            save_checkpoint(latest_path, step, epoch + 1, loss_meter.avg, B, world_size, new_lr)
