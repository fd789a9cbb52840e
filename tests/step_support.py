"""Step-level parity helpers (GPU): run the product path on cuda and the CPU oracle on the same
seeded inputs and weights.  Used by tests/test_step_gpu.py and __graft_entry__.smoke()."""
import copy
import os

import torch

from helpers import backbone_params, build_product, rel_err, step_inputs


def hyper():
    from oracle import avjepa_oracle as O
    return dict(O.DEFAULT_HYPER)


def make_train_step(enc, pred, mixed, clip_grad=None, grad_sync=None, smooth_l1_beta=None):
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
                     clip_grad=clip_grad, warmup=hp['warmup'], mixed_precision=mixed, grad_sync=grad_sync,
                     smooth_l1_beta=smooth_l1_beta)
    return step


def to_dev(masks, dev):
    return {k: [m.to(dev) for m in v] for k, v in masks.items()}


def product_forward_backward(step, clips, asgram, masks):
    """Forward + backward only (no optimizer): returns loss and a {oracle-style name: grad} dict."""
    step.keep_zh = True
    loss, _, _ = step.forward_loss(clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'])
    z, h = step.last_zh
    step.keep_zh, step.last_zh = False, None
    loss.backward()
    grads = {}
    for tag, m in (('enc', step.encoder), ('pred', step.predictor)):
        for n, p in m.named_parameters():
            if p.grad is not None and p.requires_grad:
                grads[tag + '.' + n[len('backbone.'):]] = p.grad.detach().float().cpu().clone()
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
    # bf16 (the tcgen05 / TMA production kernels) first, so the driver's launch list names them; then fp32 check mode
    for mixed, tol_loss, tol_grad in ((True, 1e-2, 5e-2), (False, 1e-4, 1e-3)):
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


FAMILIES = ('patch_embed', 'attn.qkv', 'attn.proj', 'mlp.fc1', 'mlp.fc2', 'norm', 'predictor_embed', 'mask_tokens',
            'predictor_proj')


def family_errors(grads, ref):
    """{family: (relative error over the family's tensors taken together, worst tensor name, its error)}."""
    out = {}
    for fam in FAMILIES:
        keys = [k for k in ref if fam in k]
        if not keys:
            continue
        num = sum(float(((grads[k].double() - ref[k].double()) ** 2).sum()) for k in keys)
        den = sum(float((ref[k].double() ** 2).sum()) for k in keys)
        worst = max(keys, key=lambda k: rel_err(grads[k], ref[k]) if float(ref[k].norm()) > 0 else 0.0)
        out[fam] = ((num / max(den, 1e-300)) ** 0.5, worst, rel_err(grads[worst], ref[worst]))
    return out


def autocast_oracle_grads(enc, pred, heads, clips, asgram, md, dev):
    """The reference-style bf16 arithmetic on the same GPU: the oracle's functional step under torch.autocast."""
    from oracle import avjepa_oracle as O
    ost = O.StepState({k: v.to(dev) for k, v in backbone_params(enc).items()},
                      {k: v.to(dev) for k, v in backbone_params(pred).items()}, heads=heads)
    le = {k: v.clone().requires_grad_(k not in ost.frozen) for k, v in ost.enc.items()}
    lp = {k: v.clone().requires_grad_(k not in ost.frozen) for k, v in ost.pred.items()}
    with torch.autocast('cuda', dtype=torch.bfloat16):
        lo, _, _, _, _ = O.forward_loss(ost, clips.to(dev), asgram.to(dev), md['ev'], md['ea'], md['pv'], md['pa'],
                                        enc=le, pred=lp)
    lo.backward()
    out = {}
    for tag, leaves in (('enc', le), ('pred', lp)):
        for k, t in leaves.items():
            if t.grad is not None:
                out[tag + '.' + k] = t.grad.float().cpu()
    return float(lo), out


def config_parity(model_name, heads, dev, pred_depth=12, mixed=True, seed=0):
    """Forward + backward of `model_name` (B=2, the golden masks) on the product path vs the CPU oracle (fp32) and vs
    the oracle under bf16 autocast on the GPU.  Returns a dict of the numbers the tests / bench.py assert on."""
    from oracle import avjepa_oracle as O
    torch.set_num_threads(os.cpu_count())
    clips, asgram, masks, _ = step_inputs()
    enc, pred = build_product(model_name, seed=seed, device=dev, pred_depth=pred_depth)
    st = oracle_state(enc, pred, heads)
    step = make_train_step(enc, pred, mixed)
    md = to_dev(masks, dev)
    loss, grads, z, h = product_forward_backward(step, clips.to(dev), asgram.to(dev), md)
    o = O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], keep_grads=True)
    g_err, worst, wname = grad_errors(grads, o['grads'])
    res = dict(model=model_name, batch=2, pred_depth=pred_depth, mixed=mixed, loss=loss, oracle_loss=o['loss'],
               loss_rel=abs(loss - o['loss']) / abs(o['loss']), grad_rel=g_err, worst_tensor=wname, worst_rel=worst,
               z_rel=rel_err(z[0], o['z'][0]), h_rel=rel_err(h[0], o['h'][0]),
               families={k: dict(rel=v[0], worst=v[1], worst_rel=v[2]) for k, v in family_errors(grads, o['grads']).items()})
    if mixed:
        e2, p2 = build_product(model_name, seed=seed, pred_depth=pred_depth)
        a_loss, a_grads = autocast_oracle_grads(e2, p2, heads, clips, asgram, md, dev)
        a_err, _, _ = grad_errors(a_grads, o['grads'])
        res.update(autocast_loss=a_loss, autocast_grad_rel=a_err)
    return res
