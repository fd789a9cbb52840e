"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, which exists only in the build container).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py

Fixtures (all small):
  masks_av.npz     AVMaskCollator outputs for a grid of (global seed, batch size, call index)
  masks_video.npz  MaskCollator outputs, same grid
  init_tiny.npz    per-parameter checksums of init_audio_video_model(vit_tiny) under seed 0
  init_video_tiny.npz  the same for the video-only init_video_model(vit_tiny, predictor depth 6)
  step_tiny.npz    one restated train_step (app/avjepa/train.py:437-537) on ViT-tiny, B=2,
                   seeded synthetic inputs: losses, every parameter-gradient norm, samples of
                   gradients, post-step parameter / target checksums
  posemb.npz       samples of the sincos tables
  ref_checkpoint_micro.pth.tar   a checkpoint WRITTEN BY THE REFERENCE's code path (nn.DataParallel wrappers ->
                   `module.backbone.*` keys, torch.optim.AdamW state with two steps taken, GradScaler state) for a
                   micro model (embed 16, depth 1; 32x32x4 clips) -- the wire-format fixture of
                   tests/test_checkpoint_host.py
  tensors_misc.npz repeat_interleave_batch outputs (src/utils/tensors.py:65-71)
The step is restated around the reference's own modules because the reference keeps it in a
closure that cannot be imported.
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get('AVJ_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

import logging  # noqa: E402
logging.disable(logging.CRITICAL)

from app.avjepa.utils import init_audio_video_model, init_opt  # noqa: E402
from src.masks.avmultiblock3d import AVMaskCollator  # noqa: E402
from src.masks.multiblock3d import MaskCollator  # noqa: E402
from src.masks.utils import apply_masks  # noqa: E402
from src.models.utils import pos_embs  # noqa: E402

MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None),
]


def fake_batch(b):
    return [([torch.zeros(1)], 0, [0], torch.zeros(1)) for _ in range(b)]


def golden_masks():
    av, vid = {}, {}
    for seed in (0, 234):
        for bsz in (1, 2, 8):
            torch.manual_seed(seed)
            coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
            for call in range(3):
                try:
                    _, ev, ea, pv, pa = coll(fake_batch(bsz))
                except TypeError:
                    av[f's{seed}_b{bsz}_c{call}_crash'] = np.array([1])
                    continue
                for gi in range(2):
                    for nm, t in (('ev', ev), ('ea', ea), ('pv', pv), ('pa', pa)):
                        av[f's{seed}_b{bsz}_c{call}_g{gi}_{nm}'] = t[gi].numpy().astype(np.int16)
            torch.manual_seed(seed)
            coll = MaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
            for call in range(3):
                _, e, p = coll(fake_batch(bsz))
                for gi in range(2):
                    vid[f's{seed}_b{bsz}_c{call}_g{gi}_e'] = e[gi].numpy().astype(np.int16)
                    vid[f's{seed}_b{bsz}_c{call}_g{gi}_p'] = p[gi].numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, 'masks_av.npz'), **av)
    np.savez_compressed(os.path.join(OUT, 'masks_video.npz'), **vid)


def checksum(t):
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 7 + 1
    return np.array([float(t.sum()), float(t.abs().sum()), float((t * w).sum())])


def build(model_name='vit_tiny', seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    enc, pred = init_audio_video_model(
        device=torch.device('cpu'), patch_size=16, num_frames=16, tubelet_size=2, model_name=model_name, crop_size=224,
        pred_depth=12, pred_embed_dim=384, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2,
        zero_init_mask_tokens=True, use_sdpa=True)
    return enc, pred


def golden_init():
    enc, pred = build()
    d = {}
    for tag, m in (('enc', enc), ('pred', pred)):
        for n, p in m.named_parameters():
            d[f'{tag}.{n}'] = checksum(p)
    np.savez_compressed(os.path.join(OUT, 'init_tiny.npz'), **d)


HYPER = dict(loss_exp=1.0, reg_coeff=0.0, ipe=300, ipe_scale=1.25, epochs=300, warmup=40, start_lr=0.0002, lr=0.000625,
             final_lr=1.0e-6, weight_decay=0.04, final_weight_decay=0.4, ema=(0.998, 1.0))


def reference_step(enc, pred, tgt, opt, sched, wd_sched, mom, clips, asgram, ev, ea, pv, pa):
    """app/avjepa/train.py:437-537 restated around the reference modules (fp32 CPU)."""
    new_lr = sched.step()
    new_wd = wd_sched.step()
    with torch.no_grad():
        h = tgt(clips, asgram)
        h = F.layer_norm(h, (h.size(-1),))
        vt, at = torch.split(h, [1568, 96], dim=1)
        h_v = apply_masks(vt, pv, concat=False)
        h_a = apply_masks(at, pa, concat=False)
        h = [torch.cat([a, b], dim=1) for a, b in zip(h_v, h_a)]
    masks_enc = list(zip(ev, ea))
    masks_pred = list(zip(pv, pa))
    z = enc(clips, asgram, masks_enc)
    z_t = []
    for zi, (mv, ma) in zip(z, masks_enc):
        z_t.append(torch.split(zi, [mv.shape[1], ma.shape[1]], dim=1))
    z = pred(z_t, list(zip(h_v, h_a)), masks_enc, masks_pred)
    loss_jepa = 0.
    for zi, hi in zip(z, h):
        loss_jepa += torch.mean(torch.abs(zi - hi) ** HYPER['loss_exp']) / HYPER['loss_exp']
    loss_jepa /= len(pv)
    pstd = sum([torch.sqrt(zi.var(dim=1) + 0.0001) for zi in z]) / len(z)
    loss_reg = torch.mean(F.relu(1. - pstd))
    loss = loss_jepa + HYPER['reg_coeff'] * loss_reg
    loss.backward()
    grads = {}
    for tag, m in (('enc', enc), ('pred', pred)):
        for n, p in m.named_parameters():
            if p.grad is not None:
                grads[f'{tag}.{n}'] = p.grad.detach().clone()
    opt.step()
    opt.zero_grad()
    m_ = next(mom)
    with torch.no_grad():
        for pq, pk in zip(enc.parameters(), tgt.parameters()):
            pk.data.mul_(m_).add_((1. - m_) * pq.detach().data)
    return dict(loss=float(loss), loss_jepa=float(loss_jepa), loss_reg=float(loss_reg), lr=new_lr, wd=new_wd, m=m_,
                grads=grads, z=[t.detach() for t in z], h=h)


def golden_step():
    enc, pred = build()
    tgt = copy.deepcopy(enc)
    for p in tgt.parameters():
        p.requires_grad = False
    opt, _, sched, wd_sched = init_opt(
        encoder=enc, predictor=pred, wd=HYPER['weight_decay'], final_wd=HYPER['final_weight_decay'],
        start_lr=HYPER['start_lr'], ref_lr=HYPER['lr'], final_lr=HYPER['final_lr'], iterations_per_epoch=HYPER['ipe'],
        warmup=HYPER['warmup'], num_epochs=HYPER['epochs'], ipe_scale=HYPER['ipe_scale'], mixed_precision=False)
    ema = HYPER['ema']
    n = HYPER['ipe'] * HYPER['epochs'] * HYPER['ipe_scale']
    mom = (ema[0] + i * (ema[1] - ema[0]) / n for i in range(int(n) + 1))
    torch.manual_seed(234)
    coll = AVMaskCollator(cfgs_mask=MASK_CFG, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    _, ev, ea, pv, pa = coll(fake_batch(2))
    g = torch.Generator().manual_seed(1234)
    clips = torch.randn(2, 3, 16, 224, 224, generator=g)
    asgram = -80.0 * torch.rand(2, 1, 128, 192, generator=g)
    d = {}
    for it in range(2):
        r = reference_step(enc, pred, tgt, opt, sched, wd_sched, mom, clips, asgram, ev, ea, pv, pa)
        d[f'it{it}_scalars'] = np.array([r['loss'], r['loss_jepa'], r['loss_reg'], r['lr'], r['wd'], r['m']])
        names = sorted(r['grads'])
        d[f'it{it}_grad_names'] = np.array(names)
        d[f'it{it}_grad_norms'] = np.array([float(r['grads'][k].double().norm()) for k in names])
        for k in ('enc.backbone.blocks.0.attn.qkv.weight', 'enc.backbone.patch_embed.proj.weight',
                  'pred.backbone.predictor_blocks.11.mlp.fc2.weight', 'pred.backbone.mask_tokens_v.0',
                  'enc.backbone.blocks.11.norm2.weight', 'pred.backbone.predictor_embed_a.bias'):
            d[f'it{it}_grad_sample.{k}'] = r['grads'][k].flatten()[:256].numpy()
        d[f'it{it}_z0_sample'] = r['z'][0].flatten()[:512].numpy()
        d[f'it{it}_h0_sample'] = r['h'][0].flatten()[:512].numpy()
        for tag, m in (('enc', enc), ('pred', pred), ('tgt', tgt)):
            d[f'it{it}_post.{tag}'] = np.stack([checksum(p) for _, p in m.named_parameters()])
    for gi in range(2):
        for nm, t in (('ev', ev), ('ea', ea), ('pv', pv), ('pa', pa)):
            d[f'mask_g{gi}_{nm}'] = t[gi].numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, 'step_tiny.npz'), **d)


def golden_init_video():
    """init_video_model (app/vjepa/utils.py:86-153) of the video-only twin, vit_tiny, predictor depth 6, seed 0."""
    from app.vjepa.utils import init_video_model
    torch.manual_seed(0)
    np.random.seed(0)
    enc, pred = init_video_model(
        device=torch.device('cpu'), patch_size=16, num_frames=16, tubelet_size=2, model_name='vit_tiny', crop_size=224,
        pred_depth=6, pred_embed_dim=384, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2,
        zero_init_mask_tokens=True, use_sdpa=True)
    d = {}
    for tag, m in (('enc', enc), ('pred', pred)):
        for n, p in m.named_parameters():
            d[f'{tag}.{n}'] = checksum(p)
    np.savez_compressed(os.path.join(OUT, 'init_video_tiny.npz'), **d)


def golden_posemb():
    d = {
        'v3d_1024_up': pos_embs.get_3d_sincos_pos_embed(1024, 14, 8, uniform_power=True)[::97, ::13].astype(np.float32),
        'v3d_192': pos_embs.get_3d_sincos_pos_embed(192, 14, 8, uniform_power=False)[::97, ::7].astype(np.float32),
        'a2d_384': pos_embs.get_2d_sincos_pos_embed_xy(384, 8, 12)[::5, ::11].astype(np.float32),
        'i2d_768': pos_embs.get_2d_sincos_pos_embed(768, 14)[::9, ::17].astype(np.float32),
    }
    np.savez_compressed(os.path.join(OUT, 'posemb.npz'), **d)


MICRO = dict(img_size=32, patch_size=16, num_frames=4, tubelet_size=2, embed_dim=16, depth=1, num_heads=2)


def golden_checkpoint():
    """What app/avjepa/train.py:298-300,332-350 writes, for a micro AV model: DataParallel-wrapped MultiMask wrappers,
    the 4-group AdamW of init_opt after two optimizer steps, the (bf16) GradScaler state."""
    from functools import partial
    import torch.nn as nn
    from src.models.audiovision_transformer import AudioVisionTransformer
    from src.models.audiovisionpredictor import AudioVisionTransformerPredictor
    from src.models.utils.multimask import AudioVideoMultiMaskWrapper, PredictorMultiMaskWrapper
    torch.manual_seed(7)
    ln = partial(nn.LayerNorm, eps=1e-6)
    enc = AudioVideoMultiMaskWrapper(AudioVisionTransformer(mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, **MICRO))
    pred = PredictorMultiMaskWrapper(AudioVisionTransformerPredictor(
        img_size=32, patch_size=16, num_frames=4, tubelet_size=2, embed_dim=16, predictor_embed_dim=8, depth=1, num_heads=2,
        mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2))
    tgt = copy.deepcopy(enc)
    opt, scaler, sched, wd_sched = init_opt(
        encoder=enc, predictor=pred, wd=0.04, final_wd=0.4, start_lr=2e-4, ref_lr=6.25e-4, final_lr=1e-6,
        iterations_per_epoch=3, warmup=1, num_epochs=2, ipe_scale=1.25, mixed_precision=True)
    enc, pred, tgt = (torch.nn.DataParallel(m) for m in (enc, pred, tgt))
    g = torch.Generator().manual_seed(5)
    for _ in range(2):
        sched.step()
        wd_sched.step()
        for p in list(enc.parameters()) + list(pred.parameters()):
            if p.requires_grad:
                p.grad = torch.randn(p.shape, generator=g) * 1e-2
        opt.step()
        opt.zero_grad()
    save_dict = {
        'encoder': enc.state_dict(), 'predictor': pred.state_dict(), 'opt': opt.state_dict(),
        'scaler': None if scaler is None else scaler.state_dict(), 'target_encoder': tgt.state_dict(),
        'epoch': 1, 'loss': 0.5, 'batch_size': 2, 'world_size': 1, 'lr': 6.25e-4,
    }
    torch.save(save_dict, os.path.join(OUT, 'ref_checkpoint_micro.pth.tar'))


def golden_misc():
    from src.utils.tensors import repeat_interleave_batch
    x = torch.arange(6 * 3, dtype=torch.float32).reshape(6, 3)
    d = {'rib_in': x.numpy()}
    for B, rep in ((2, 1), (2, 2), (3, 2), (1, 3), (6, 2)):
        d[f'rib_B{B}_r{rep}'] = repeat_interleave_batch(x, B, rep).numpy()
    np.savez_compressed(os.path.join(OUT, 'tensors_misc.npz'), **d)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    golden_posemb()
    golden_masks()
    golden_init()
    golden_init_video()
    golden_step()
    golden_checkpoint()
    golden_misc()
    print('golden fixtures written to', OUT, 'torch', torch.__version__)
