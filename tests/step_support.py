"""Step-level parity helpers (GPU): run the product path on cuda and the CPU oracle on the same
seeded inputs and weights.  Used by tests/test_step_gpu.py and __graft_entry__.smoke()."""
import copy
import os

import torch

from helpers import backbone_params, build_product, rel_err, step_inputs


def hyper():
    from oracle import avjepa_oracle as O
    return dict(O.DEFAULT_HYPER)


def make_train_step(enc, pred, mixed, clip_grad=None):
    from avjepa_b200.app.avjepa.train import TrainStep
    from avjepa_b200.app.avjepa.utils import init_opt
    hp = hyper()
    tgt = copy.deepcopy(enc)
    opt, scaler, sched, wd_sched = init_opt(
        encoder=enc, predictor=pred, wd=hp['weight_decay'], final_wd=hp['final_weight_decay'], start_lr=hp['start_lr'],
        ref_lr=hp['lr'], final_lr=hp['final_lr'], iterations_per_epoch=hp['ipe'], warmup=hp['warmup'],
        num_epochs=hp['epochs'], ipe_scale=hp['ipe_scale'], mixed_precision=mixed)
    n = hp['ipe'] * hp['epochs'] * hp['ipe_scale']
    mom = (hp['ema'][0] + i * (hp['ema'][1] - hp['ema'][0]) / n for i in range(int(n) + 1))
    step = TrainStep(enc, pred, tgt, opt, scaler, sched, wd_sched, mom, loss_exp=hp['loss_exp'], reg_coeff=hp['reg_coeff'],
                     clip_grad=clip_grad, warmup=hp['warmup'], mixed_precision=mixed)
    return step


def to_dev(masks, dev):
    return {k: [m.to(dev) for m in v] for k, v in masks.items()}


def product_forward_backward(step, clips, asgram, masks):
    """Forward + backward only (no optimizer): returns loss and a {oracle-style name: grad} dict."""
    from avjepa_b200 import loss as L
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=step.mixed_precision):
        h = step.forward_target(clips, asgram, masks['pv'], masks['pa'])
        z = step.forward_context(clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'])
        loss = L.jepa_loss(z, h, 1.0)
    loss.backward()
    grads = {}
    for tag, m in (('enc', step.encoder), ('pred', step.predictor)):
        for n, p in m.named_parameters():
            if p.grad is not None and p.requires_grad:
                grads[tag + '.' + n[len('backbone.'):]] = p.grad.detach().float().cpu().clone()
    step.optimizer.mark_grads_dirty()
    step.optimizer.zero_grad()
    return float(loss), grads, [t.detach().float().cpu() for t in z], [t.float().cpu() for t in h]


def grad_errors(grads, ref):
    """(global relative error, worst per-tensor relative error among non-negligible tensors, name)."""
    assert set(grads) == set(ref), set(grads) ^ set(ref)
    num = sum(float(((grads[k].double() - ref[k].double()) ** 2).sum()) for k in ref)
    den = sum(float((ref[k].double() ** 2).sum()) for k in ref)
    gnorm = den ** 0.5
    worst, wname = 0.0, None
    for k in ref:
        if float(ref[k].double().norm()) < 1e-3 * gnorm / max(1, len(ref)) ** 0.5:
            continue
        e = rel_err(grads[k], ref[k])
        if e > worst:
            worst, wname = e, k
    return (num / den) ** 0.5, worst, wname


def oracle_state(enc, pred, heads):
    from oracle import avjepa_oracle as O
    return O.StepState(backbone_params(enc), backbone_params(pred), heads=heads)


def build_tiny_step(dev, seed=0, mixed=True):
    """(TrainStep on a seeded ViT-tiny / predictor-depth-2 pair, positional batch tuple for step(*batch))."""
    clips, asgram, masks, _ = step_inputs()
    enc, pred = build_product('vit_tiny', seed=seed, device=dev, pred_depth=2)
    step = make_train_step(enc, pred, mixed)
    md = to_dev(masks, dev)
    return step, (clips.to(dev), asgram.to(dev), md['ev'], md['ea'], md['pv'], md['pa'])


def run_smoke(dev):
    """__graft_entry__.smoke(): tiny step on cuda:0 in both modes vs the live CPU oracle."""
    from oracle import avjepa_oracle as O
    torch.set_num_threads(os.cpu_count())
    clips, asgram, masks, _ = step_inputs()
    for mixed, tol_loss, tol_grad in ((False, 1e-4, 1e-3), (True, 1e-2, 5e-2)):
        enc, pred = build_product('vit_tiny', seed=0, device=dev, pred_depth=2)
        st = oracle_state(enc, pred, 3)
        step = make_train_step(enc, pred, mixed)
        loss, grads, z, h = product_forward_backward(step, clips.to(dev), asgram.to(dev), to_dev(masks, dev))
        o = O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], keep_grads=True)
        g_err, worst, wname = grad_errors(grads, o['grads'])
        print(f'smoke mixed={mixed}: loss {loss:.6f} vs oracle {o["loss"]:.6f}; grad rel err {g_err:.2e} '
              f'(worst tensor {wname}: {worst:.2e})')
        assert abs(loss - o['loss']) <= tol_loss * abs(o['loss']), (loss, o['loss'])
        assert g_err <= tol_grad, g_err
        # and one full fused step (AdamW + EMA) must run
        out = step(clips.to(dev), asgram.to(dev), *[to_dev(masks, dev)[k] for k in ('ev', 'ea', 'pv', 'pa')])
        assert abs(out[0] - o['loss']) <= tol_loss * abs(o['loss'])
    torch.cuda.synchronize()
    print('smoke ok')
