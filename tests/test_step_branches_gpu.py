"""Branches of the step that the ViT-tiny/epoch-0 parity tests never reach: gradient clipping (``epoch > warmup``,
``app/avjepa/train.py:518-520``), the smooth-L1 loss mode, device-side ``grad_logger`` / ``adamw_logger`` statistics,
the reference's own call order around the drop-in optimizer / scaler (``train.py:514-531``), and optimizer-state
interchange with ``torch.optim.AdamW``."""
import copy
import os

import numpy as np
import pytest
import torch

from helpers import backbone_params, build_product, rel_err, step_inputs
import step_support as S

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _oracle(enc, pred, **hyper):
    from oracle import avjepa_oracle as O
    torch.set_num_threads(os.cpu_count())
    hp = dict(O.DEFAULT_HYPER)
    hp.update(hyper)
    return O, O.StepState(backbone_params(enc), backbone_params(pred), heads=3), hp


def _post_step_close(step, st, lr, what):
    for tag, module, ref in (('enc', step.encoder, st.enc), ('pred', step.predictor, st.pred), ('tgt', step.target_encoder, st.tgt)):
        for n, p in module.named_parameters():
            d = (p.detach().float().cpu() - ref[n[len('backbone.'):]]).abs()
            assert float(d.max()) <= 2.5 * lr, (what, tag, n, float(d.max()))
            assert float(d.mean()) <= 0.05 * lr, (what, tag, n, float(d.mean()))


def test_clip_branch_matches_oracle():
    """epoch = warmup + 1 and a clip threshold far below the gradient norms: both norms and the clipped update."""
    clips, asgram, masks, _ = step_inputs()
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    clip = 0.002                        # the tiny model's gradient norms are ~7e-3 (encoder) and larger (predictor)
    O, st, hp = _oracle(enc, pred, clip_grad=clip)
    step = S.make_train_step(enc, pred, mixed=False, clip_grad=clip)
    md = S.to_dev(masks, DEV)
    epoch = hp['warmup'] + 1
    out = step(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'], epoch=epoch)
    o = O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], hyper=hp, epoch=epoch)
    assert o['enc_grad_norm'] > 2 * clip and o['pred_grad_norm'] > 2 * clip      # the threshold really bites
    assert out[0] == pytest.approx(o['loss'], rel=1e-4)
    assert step.last['enc_norm'] == pytest.approx(o['enc_grad_norm'], rel=1e-4)
    assert step.last['pred_norm'] == pytest.approx(o['pred_grad_norm'], rel=1e-4)
    _post_step_close(step, st, o['lr'], 'clip')
    # and with the clip disabled the two runs must DIFFER from the clipped ones in Adam's second moment (sanity of the test)
    p = dict(step.predictor.named_parameters())['backbone.predictor_proj.weight']
    v = step.optimizer.state[p]['exp_avg_sq']
    g2 = 0.001 * (o['pred_grad_norm'] ** 2) * (clip / o['pred_grad_norm']) ** 2      # (1-beta2) * |clipped g|^2 upper bound
    assert float(v.sum()) <= g2 * 1.001


def test_smooth_l1_mode_matches_oracle():
    clips, asgram, masks, _ = step_inputs()
    for mixed, tol_l, tol_g in ((False, 1e-4, 1e-4), (True, 1e-2, 5e-2)):
        enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
        O, st, hp = _oracle(enc, pred, smooth_l1_beta=0.5)
        step = S.make_train_step(enc, pred, mixed=mixed, smooth_l1_beta=0.5)
        loss, grads, z, h = S.product_forward_backward(step, clips.to(DEV), asgram.to(DEV), S.to_dev(masks, DEV))
        o = O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], hyper=hp, keep_grads=True)
        g_err, worst, wname = S.grad_errors(grads, o['grads'])
        assert abs(loss - o['loss']) <= tol_l * abs(o['loss']), (mixed, loss, o['loss'])
        assert g_err <= tol_g, (mixed, g_err, wname, worst)
        assert abs(o['loss'] - 0.856063) > 1e-2                    # the mode really differs from L1 (0.856063) on these inputs


def test_device_side_logging_statistics_match_torch():
    """TrainStep(log_stats=True): grad_logger / adamw_logger numbers from ONE segment-reduction pass per flat buffer
    equal the per-tensor torch norms of the same gradients / moments (src/utils/logging.py:91-118)."""
    from avjepa_b200.src.utils.logging import adamw_logger, grad_logger
    clips, asgram, masks, _ = step_inputs()
    md = S.to_dev(masks, DEV)
    args = (clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    step = S.make_train_step(enc, pred, mixed=True)
    step.optimizer.ensure_built()
    # reference numbers: forward+backward, per-tensor torch norms, then throw the gradients away
    loss, _, _ = step.forward_loss(*args)
    loss.backward()
    ref_enc = grad_logger(step.encoder.named_parameters())
    ref_pred = grad_logger(step.predictor.named_parameters())
    assert ref_enc.count > 0 and ref_enc.first_layer > 0 and ref_pred.last_layer > 0
    step.optimizer.zero_grad()
    out = step(*args, log_stats=True)
    g_enc, g_pred, optim = step.stats
    for a, b in ((g_enc, ref_enc), (g_pred, ref_pred)):
        assert a.count == b.count
        for f in ('avg', 'min', 'max', 'first_layer', 'last_layer'):
            assert getattr(a, f) == pytest.approx(getattr(b, f), rel=2e-2), f     # bf16 run-to-run: split-K atomics
    ref_opt = adamw_logger_torch(step.optimizer)
    for k in ('exp_avg', 'exp_avg_sq'):
        assert optim[k].count == ref_opt[k].count
        for f in ('avg', 'min', 'max'):
            assert getattr(optim[k], f) == pytest.approx(getattr(ref_opt[k], f), rel=1e-5), (k, f)
    # the public adamw_logger goes through the same kernel for the fused optimizer
    pub = adamw_logger(step.optimizer)
    assert pub['exp_avg'].avg == pytest.approx(ref_opt['exp_avg'].avg, rel=1e-5)
    assert np.isfinite(out[0])


def adamw_logger_torch(optimizer):
    from avjepa_b200.src.utils.logging import AverageMeter
    a, b = AverageMeter(), AverageMeter()
    for g in optimizer.param_groups:
        for p in g['params']:
            st = optimizer.state.get(p)
            if st and 'exp_avg' in st:
                a.update(float(st['exp_avg'].abs().mean()))
                b.update(float(st['exp_avg_sq'].abs().mean()))
    return {'exp_avg': a, 'exp_avg_sq': b}


def test_reference_call_order_around_dropin_optimizer():
    """The reference loop's own sequence (train.py:514-531) with the drop-in optimizer / scaler:
    scale -> backward -> unscale_ -> clip_grad_norm_ -> scaler.step -> grad_logger -> zero_grad -> adamw_logger -> EMA.
    Clipping must see true (unscaled) gradients, grad_logger must see non-zero gradients after the step, and the
    resulting parameters must track the oracle's clipped update."""
    from avjepa_b200.src.utils.logging import adamw_logger, grad_logger
    clips, asgram, masks, _ = step_inputs()
    md = S.to_dev(masks, DEV)
    clip = 0.002
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    O, st, hp = _oracle(enc, pred, clip_grad=clip)
    step = S.make_train_step(enc, pred, mixed=True, clip_grad=clip)       # only used as a container of the pieces
    opt, scaler = step.optimizer, step.scaler
    assert scaler is not None and scaler.get_scale() == 1.0
    new_lr, new_wd = step.scheduler.step(), step.wd_scheduler.step()
    loss, _, _ = step.forward_loss(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    enc_norm = torch.nn.utils.clip_grad_norm_(step.encoder.parameters(), clip)
    pred_norm = torch.nn.utils.clip_grad_norm_(step.predictor.parameters(), clip)
    scaler.step(opt)
    scaler.update()
    gs = grad_logger(step.encoder.named_parameters())
    assert gs.count > 0 and gs.max > 0.0 and gs.first_layer > 0.0                      # not zeroed by the step
    opt.zero_grad()
    assert all(float(p.grad.abs().max()) == 0.0 for p in step.encoder.parameters() if p.grad is not None)
    os_ = adamw_logger(opt)
    assert os_['exp_avg'].count > 0 and os_['exp_avg'].max > 0
    m = next(step.momentum_scheduler)
    with torch.no_grad():
        for pq, pk in zip(step.encoder.parameters(), step.target_encoder.parameters()):
            pk.data.mul_(m).add_((1. - m) * pq.detach().data)
    o = O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], hyper=hp, epoch=hp['warmup'] + 1)
    assert float(enc_norm) == pytest.approx(o['enc_grad_norm'], rel=2e-2)
    assert float(pred_norm) == pytest.approx(o['pred_grad_norm'], rel=2e-2)
    for tag, module, ref in (('enc', step.encoder, st.enc), ('pred', step.predictor, st.pred), ('tgt', step.target_encoder, st.tgt)):
        for n, p in module.named_parameters():
            d = (p.detach().float().cpu() - ref[n[len('backbone.'):]]).abs()
            assert float(d.mean()) <= 0.1 * o['lr'], (tag, n, float(d.mean()))
    # a second iteration through the same order must not see stale gradients (zero_grad really zeroed)
    loss2, _, _ = step.forward_loss(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    scaler.scale(loss2).backward()
    g2 = float(torch.nn.utils.clip_grad_norm_(step.encoder.parameters(), 1e9))
    assert 0.1 * float(enc_norm) < g2 < 2.0 * float(enc_norm)       # one Adam step at init moves the norm, stale grads would add to it
    # a skipped step (NaN guard / accumulation): zero_grad() must still clear what backward wrote
    opt.zero_grad()
    assert float(opt.grad_norm_sq()) == 0.0


def test_optimizer_state_loads_into_torch_adamw_and_back():
    """state_dict() of the fused optimizer has one independent `step` per parameter and plain moment tensors: it
    loads into torch.optim.AdamW (the reference's optimizer), which then steps with the right bias correction; and
    a torch.optim.AdamW state dict loads into the fused optimizer."""
    clips, asgram, masks, _ = step_inputs()
    md = S.to_dev(masks, DEV)
    args = (clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    step = S.make_train_step(enc, pred, mixed=False)
    step(*args)
    step(*args)
    sd = step.optimizer.state_dict()
    steps = [st['step'] for st in sd['state'].values()]
    assert len({id(s) for s in steps}) == len(steps) and all(float(s) == 2.0 for s in steps)
    # clone the models into a torch.optim.AdamW with the same group structure
    enc2, pred2 = copy.deepcopy(step.encoder), copy.deepcopy(step.predictor)
    groups = []
    for g, (m2,) in zip(step.optimizer.param_groups, ((enc2,), (pred2,), (enc2,), (pred2,))):
        own = {id(p) for p in g['params']}
        src = step.encoder if m2 is enc2 else step.predictor
        names = [n for n, p in src.named_parameters() if id(p) in own]
        d2 = dict(m2.named_parameters())
        groups.append({'params': [d2[n] for n in names], 'weight_decay': g['weight_decay'], 'lr': g['lr']})
    ref_opt = torch.optim.AdamW(groups, betas=(0.9, 0.999), eps=1e-8)
    ref_opt.load_state_dict(sd)
    for g in ref_opt.param_groups:
        for p in g['params']:
            if p.requires_grad:
                p.grad = torch.full_like(p, 1e-3)
    ref_opt.step()
    some = [p for g in ref_opt.param_groups for p in g['params'] if p in ref_opt.state][0]
    assert float(ref_opt.state[some]['step']) == 3.0                 # not 3 x (number of parameters)
    # and back: a stock AdamW state dict into a fresh fused optimizer
    enc3, pred3 = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    step3 = S.make_train_step(enc3, pred3, mixed=False)
    step3.optimizer.load_state_dict(ref_opt.state_dict())
    step3.optimizer.ensure_built()
    p3 = dict(step3.predictor.named_parameters())['backbone.predictor_proj.weight']
    p2 = dict(pred2.named_parameters())['backbone.predictor_proj.weight']
    assert rel_err(step3.optimizer.state[p3]['exp_avg'], ref_opt.state[p2]['exp_avg']) < 1e-6
    assert step3.optimizer._step == 3
    out = step3(*args)
    assert np.isfinite(out[0])


def test_merged_mask_schedule_equals_the_per_mask_loop(monkeypatch):
    """The MultiMask wrappers run all masks through ONE variable-length stack (default) -- the reference loops over the
    masks.  Same predictions (row-wise identical arithmetic) and same gradients (sums in a different order)."""
    clips, asgram, masks, _ = step_inputs()
    md = S.to_dev(masks, DEV)
    res = {}
    for merged in ('1', '0'):
        monkeypatch.setenv('AVJ_MERGE_MASKS', merged)
        enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
        step = S.make_train_step(enc, pred, mixed=False)
        res[merged] = S.product_forward_backward(step, clips.to(DEV), asgram.to(DEV), md)
    (l1, g1, z1, h1), (l0, g0, z0, h0) = res['1'], res['0']
    assert l1 == pytest.approx(l0, rel=1e-6)
    for a, b in zip(z1, z0):
        assert torch.equal(a, b)
    g_err, worst, wname = S.grad_errors(g1, g0)
    assert g_err < 1e-5 and worst < 1e-4, (g_err, wname, worst)
    # video-only wrappers take the same route
    import avjepa_b200.src.models.vision_transformer as vit
    from avjepa_b200.src.models.predictor import vit_predictor
    from avjepa_b200.src.models.utils.multimask import MultiMaskWrapper, PredictorMultiMaskWrapper
    torch.manual_seed(1)
    venc = MultiMaskWrapper(vit.vit_tiny(img_size=224, num_frames=16, tubelet_size=2, uniform_power=True)).to(DEV)
    vpred = PredictorMultiMaskWrapper(vit_predictor(img_size=224, num_frames=16, tubelet_size=2, embed_dim=192, predictor_embed_dim=384,
                                                    depth=2, num_heads=3, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2)).to(DEV)
    outs = {}
    for merged in ('1', '0'):
        monkeypatch.setenv('AVJ_MERGE_MASKS', merged)
        z = venc(clips.to(DEV), md['ev'])
        o = vpred(z, [None, None], md['ev'], md['pv'])
        (o[0].square().mean() + o[1].square().mean()).backward()
        outs[merged] = ([t.detach().clone() for t in o], {n: p.grad.clone() for n, p in venc.named_parameters() if p.grad is not None})
        for p in list(venc.parameters()) + list(vpred.parameters()):
            p.grad = None
    for a, b in zip(outs['1'][0], outs['0'][0]):
        assert torch.equal(a, b)
    for n in outs['1'][1]:
        assert rel_err(outs['1'][1][n], outs['0'][1][n]) < 1e-4, n
