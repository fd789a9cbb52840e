"""Step-level parity on the GPU: the product path (public module API -> C-ABI kernels) against
the CPU oracle on the same seeded inputs and bit-identical weights, and against the golden
vectors generated from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from helpers import backbone_params, build_product, checksum, golden, rel_err, step_inputs
import step_support as S

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def oracle_ref():
    """Two oracle iterations on ViT-tiny/B=2 (pred depth 12) -- the same run the golden fixture pins."""
    from oracle import avjepa_oracle as O
    torch.set_num_threads(os.cpu_count())
    enc, pred = build_product('vit_tiny', seed=0)
    st = O.StepState(backbone_params(enc), backbone_params(pred), heads=3)
    clips, asgram, masks, gz = step_inputs()
    outs = [O.train_step(st, clips, asgram, masks['ev'], masks['ea'], masks['pv'], masks['pa'], keep_grads=True)
            for _ in range(2)]
    return st, outs


def test_fp32_check_mode_matches_oracle_and_golden(oracle_ref):
    """fp32 check mode: loss and gradients within 1e-4 relative of the reference's fp32 path."""
    st, outs = oracle_ref
    clips, asgram, masks, gz = step_inputs()
    enc, pred = build_product('vit_tiny', seed=0, device=DEV)
    step = S.make_train_step(enc, pred, mixed=False)
    loss, grads, z, h = S.product_forward_backward(step, clips.to(DEV), asgram.to(DEV), S.to_dev(masks, DEV))
    assert loss == pytest.approx(gz['it0_scalars'][0], rel=1e-4)          # reference golden
    assert loss == pytest.approx(outs[0]['loss'], rel=1e-4)               # live oracle
    assert rel_err(z[0].flatten()[:512], torch.from_numpy(gz['it0_z0_sample'])) < 1e-4
    assert rel_err(h[0].flatten()[:512], torch.from_numpy(gz['it0_h0_sample'])) < 1e-4
    g_err, worst, wname = S.grad_errors(grads, outs[0]['grads'])
    assert g_err < 1e-4, g_err
    assert worst < 1e-3, (wname, worst)
    names = [str(n).replace('.backbone.', '.', 1) for n in gz['it0_grad_names']]
    for n, gn in zip(names, gz['it0_grad_norms']):
        assert float(grads[n].double().norm()) == pytest.approx(gn, rel=1e-3, abs=1e-8), n


def test_fp32_two_full_steps_track_the_oracle(oracle_ref):
    """Schedules + AdamW + EMA: second-iteration loss and post-step weights follow the oracle."""
    st, outs = oracle_ref
    clips, asgram, masks, gz = step_inputs()
    enc, pred = build_product('vit_tiny', seed=0, device=DEV)
    step = S.make_train_step(enc, pred, mixed=False)
    md = S.to_dev(masks, DEV)
    res = [step(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa']) for _ in range(2)]
    for it in range(2):
        assert res[it][0] == pytest.approx(outs[it]['loss'], rel=2e-4)
        assert res[it][3] == outs[it]['lr'] and res[it][4] == outs[it]['wd']
        assert res[it][0] == pytest.approx(gz[f'it{it}_scalars'][0], rel=2e-4)
    lr = outs[1]['lr']
    for tag, module, ref in (('enc', step.encoder, st.enc), ('pred', step.predictor, st.pred),
                             ('tgt', step.target_encoder, st.tgt)):
        for n, p in module.named_parameters():
            r = ref[n[len('backbone.'):]]
            d = (p.detach().float().cpu() - r).abs()
            # Adam's first steps are +-lr per element; sign noise on ~zero gradients bounds the
            # per-element gap by 2 steps x 2 lr, the mean gap must be far smaller
            assert float(d.max()) <= 4.5 * lr, (tag, n, float(d.max()))
            assert float(d.mean()) <= 0.05 * lr, (tag, n, float(d.mean()))


def test_bf16_matches_oracle_within_reference_bf16_error(oracle_ref):
    """bf16 production path: loss within 1e-2 of fp32; gradient error no worse than what the
    reference's own bf16-autocast arithmetic (oracle run under torch.autocast on the GPU) shows."""
    from oracle import avjepa_oracle as O
    st, outs = oracle_ref
    clips, asgram, masks, gz = step_inputs()
    enc, pred = build_product('vit_tiny', seed=0, device=DEV)
    step = S.make_train_step(enc, pred, mixed=True)
    md = S.to_dev(masks, DEV)
    loss, grads, z, h = S.product_forward_backward(step, clips.to(DEV), asgram.to(DEV), md)
    assert loss == pytest.approx(outs[0]['loss'], rel=1e-2)
    g_err, worst, wname = S.grad_errors(grads, outs[0]['grads'])
    # the reference's bf16 arithmetic on the same GPU
    e2, p2 = build_product('vit_tiny', seed=0)
    ost = O.StepState({k: v.to(DEV) for k, v in backbone_params(e2).items()},
                      {k: v.to(DEV) for k, v in backbone_params(p2).items()}, heads=3)
    le = {k: v.clone().requires_grad_(k not in ost.frozen) for k, v in ost.enc.items()}
    lp = {k: v.clone().requires_grad_(k not in ost.frozen) for k, v in ost.pred.items()}
    with torch.autocast('cuda', dtype=torch.bfloat16):
        lo, _, _, _, _ = O.forward_loss(ost, clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'],
                                        enc=le, pred=lp)
    lo.backward()
    ref_bf16 = {}
    for tag, leaves in (('enc', le), ('pred', lp)):
        for k, t in leaves.items():
            if t.grad is not None:
                ref_bf16[tag + '.' + k] = t.grad.float().cpu()
    a_err, _, _ = S.grad_errors(ref_bf16, outs[0]['grads'])
    print(f'bf16 grad rel err: ours {g_err:.3e} (worst {wname} {worst:.3e}); reference-style autocast {a_err:.3e}; '
          f'loss ours {loss:.6f} autocast {float(lo):.6f} fp32 {outs[0]["loss"]:.6f}')
    assert g_err <= max(1e-2, 1.25 * a_err), (g_err, a_err)


def test_frozen_encoder_forward_full_tokens():
    """BASELINE config 5 shape: no masks, all 1664 tokens, no grad."""
    from oracle import avjepa_oracle as O
    clips, asgram, masks, _ = step_inputs()
    enc, _ = build_product('vit_tiny', seed=0, device=DEV)
    with torch.no_grad():
        out = enc(clips.to(DEV), asgram.to(DEV))
        ref = O.av_encoder_forward(backbone_params(enc), clips, asgram, 3)
        assert out.shape == (2, 1664, 192)
        assert rel_err(out, ref) < 1e-4
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out16 = enc(clips.to(DEV), asgram.to(DEV))
        assert rel_err(out16, ref) < 2e-2


def test_empty_audio_context_and_video_only_models():
    """Edge cases: Kc_a == 0 and the video-only encoder/predictor pair."""
    from oracle import avjepa_oracle as O
    import avjepa_b200.src.models.vision_transformer as vit
    from avjepa_b200.src.models.predictor import vit_predictor
    clips, asgram, masks, _ = step_inputs()
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    ev, pv, pa = masks['ev'][1], masks['pv'][1], masks['pa'][1]
    ea = torch.zeros((2, 0), dtype=torch.int64)
    z = enc(clips.to(DEV), asgram.to(DEV), [(ev.to(DEV), ea.to(DEV))])[0]
    zr = O.av_encoder_forward(backbone_params(enc), clips, asgram, 3, masks=(ev, ea))
    assert rel_err(z, zr) < 1e-4
    out = pred([(z[:, :ev.shape[1]], z[:, ev.shape[1]:])], [None], [(ev.to(DEV), ea.to(DEV))], [(pv.to(DEV), pa.to(DEV))])[0]
    outr = O.av_predictor_forward(backbone_params(pred), zr[:, :ev.shape[1]], zr[:, ev.shape[1]:], (ev, ea), (pv, pa), 0, 3)
    assert rel_err(out, outr) < 1e-4
    out.sum().backward()
    # video-only pair
    torch.manual_seed(1)
    venc = vit.vit_tiny(img_size=224, num_frames=16, tubelet_size=2, uniform_power=True).to(DEV)
    vpred = vit_predictor(img_size=224, num_frames=16, tubelet_size=2, embed_dim=192, predictor_embed_dim=384, depth=2,
                          num_heads=3, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2).to(DEV)
    pe = {k: v.detach().cpu().float() for k, v in venc.state_dict().items()}
    pp = {k: v.detach().cpu().float() for k, v in vpred.state_dict().items()}
    zv = venc(clips.to(DEV), masks=ev.to(DEV))
    zvr = O.video_encoder_forward(pe, clips, 3, masks=ev)
    assert rel_err(zv, zvr) < 1e-4
    o = vpred(zv, None, ev.to(DEV), pv.to(DEV), mask_index=1)
    orf = O.video_predictor_forward(pp, zvr, ev, pv, 1, 3)
    assert rel_err(o, orf) < 1e-4
    le = {k: v.clone().requires_grad_('pos_embed' not in k) for k, v in pe.items()}
    lp = {k: v.clone().requires_grad_('pos_embed' not in k) for k, v in pp.items()}
    O.video_predictor_forward(lp, O.video_encoder_forward(le, clips, 3, masks=ev), ev, pv, 1, 3).square().mean().backward()
    o.square().mean().backward()
    for n, p in venc.named_parameters():
        if p.requires_grad and le[n].grad is not None and float(le[n].grad.norm()) > 1e-7:
            assert rel_err(p.grad, le[n].grad) < 2e-3, n
    for n, p in vpred.named_parameters():
        if p.requires_grad and lp[n].grad is not None and float(lp[n].grad.norm()) > 1e-7:
            assert rel_err(p.grad, lp[n].grad) < 2e-3, n


def test_round_trip_checkpoint_keys(tmp_path):
    """Checkpoint wire format: same top-level keys as the reference, loads back."""
    from avjepa_b200.app.avjepa.train import save_checkpoint
    from avjepa_b200.app.avjepa.utils import load_checkpoint
    enc, pred = build_product('vit_tiny', seed=0, device=DEV, pred_depth=2)
    step = S.make_train_step(enc, pred, mixed=True)
    clips, asgram, masks, _ = step_inputs()
    md = S.to_dev(masks, DEV)
    step(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    path = str(tmp_path / 'ck.pth.tar')
    save_checkpoint(path, step, 1, 0.5, 2, 1, 1e-4)
    ck = torch.load(path, map_location='cpu')
    assert set(ck) == {'encoder', 'predictor', 'opt', 'scaler', 'target_encoder', 'epoch', 'loss', 'batch_size',
                       'world_size', 'lr'}
    assert 'backbone.blocks.0.attn.qkv.weight' in ck['encoder']
    before = checksum(step.encoder.state_dict()['backbone.blocks.0.attn.qkv.weight'])
    e, p, t, o, s, epoch = load_checkpoint(path, step.encoder, step.predictor, step.target_encoder, step.optimizer, step.scaler)
    assert epoch == 1
    assert np.array_equal(before, checksum(e.state_dict()['backbone.blocks.0.attn.qkv.weight']))
    out = step(clips.to(DEV), asgram.to(DEV), md['ev'], md['ea'], md['pv'], md['pa'])
    assert np.isfinite(out[0])
