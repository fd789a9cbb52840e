"""Checkpoint wire format (row f2): a checkpoint WRITTEN BY THE REFERENCE's code path (tests/golden/
ref_checkpoint_micro.pth.tar, made by oracle/make_golden.py::golden_checkpoint: nn.DataParallel wrappers ->
`module.backbone.*` keys, torch.optim.AdamW state, GradScaler state) loads into this package's modules and optimizer;
failures are loud; the evals' `load_pretrained` key stripping; and the reverse direction into the live reference."""
import copy
import os
import sys
from functools import partial

import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN, REFERENCE

CK = os.path.join(GOLDEN, 'ref_checkpoint_micro.pth.tar')
MICRO = dict(img_size=32, patch_size=16, num_frames=4, tubelet_size=2, embed_dim=16, depth=1, num_heads=2)


def build_micro():
    from avjepa_b200.app.avjepa.utils import init_opt
    from avjepa_b200.src.models.audiovision_transformer import AudioVisionTransformer
    from avjepa_b200.src.models.audiovisionpredictor import AudioVisionTransformerPredictor
    from avjepa_b200.src.models.utils.multimask import AudioVideoMultiMaskWrapper, PredictorMultiMaskWrapper
    torch.manual_seed(1)
    ln = partial(nn.LayerNorm, eps=1e-6)
    enc = AudioVideoMultiMaskWrapper(AudioVisionTransformer(mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, **MICRO))
    pred = PredictorMultiMaskWrapper(AudioVisionTransformerPredictor(
        img_size=32, patch_size=16, num_frames=4, tubelet_size=2, embed_dim=16, predictor_embed_dim=8, depth=1, num_heads=2,
        mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2))
    tgt = copy.deepcopy(enc)
    opt, scaler, sched, wd_sched = init_opt(encoder=enc, predictor=pred, wd=0.04, final_wd=0.4, start_lr=2e-4, ref_lr=6.25e-4,
                                            final_lr=1e-6, iterations_per_epoch=3, warmup=1, num_epochs=2, ipe_scale=1.25,
                                            mixed_precision=True)
    return enc, pred, tgt, opt, scaler


def test_reference_written_checkpoint_loads():
    from avjepa_b200.app.avjepa.utils import load_checkpoint
    ck = torch.load(CK, map_location='cpu')
    assert all(k.startswith('module.backbone.') for k in ck['encoder'])           # it really is the reference's format
    enc, pred, tgt, opt, scaler = build_micro()
    before = enc.state_dict()['backbone.blocks.0.attn.qkv.weight'].clone()
    e, p, t, o, s, epoch = load_checkpoint(CK, enc, pred, tgt, opt, scaler)
    assert epoch == 1
    for mod, key in ((enc, 'encoder'), (pred, 'predictor'), (tgt, 'target_encoder')):
        sd = mod.state_dict()
        assert set(sd) == {k[len('module.'):] for k in ck[key]}
        for k, v in sd.items():
            assert torch.equal(v, ck[key]['module.' + k]), k
    assert not torch.equal(before, enc.state_dict()['backbone.blocks.0.attn.qkv.weight'])
    # AdamW state: same parameter indexing (4 groups, frozen sincos tables included in group 0), two steps taken
    idx = {id(p): i for i, p in enumerate(q for g in opt.param_groups for q in g['params'])}
    n_state = 0
    for p_, st in opt.state.items():
        ref = ck['opt']['state'][idx[id(p_)]]
        assert torch.equal(st['exp_avg'], ref['exp_avg']) and torch.equal(st['exp_avg_sq'], ref['exp_avg_sq'])
        n_state += 1
    assert n_state == len(ck['opt']['state']) and opt._step == 2
    assert [g['lr'] for g in opt.param_groups] == [g['lr'] for g in ck['opt']['param_groups']]
    # re-serialising gives every parameter its own step tensor
    sd = opt.state_dict()
    steps = [st['step'] for st in sd['state'].values()]
    assert len({id(x) for x in steps}) == len(steps) and all(float(x) == 2.0 for x in steps)
    assert s.get_scale() == 1.0                                                    # bf16 needs no loss scaling


def test_mismatched_checkpoint_is_an_error_not_epoch_zero(tmp_path):
    from avjepa_b200.app.avjepa.utils import load_checkpoint
    ck = torch.load(CK, map_location='cpu')
    bad = dict(ck)
    bad['encoder'] = {k.replace('blocks.0.attn.qkv', 'blocks.0.attn.qkvx'): v for k, v in ck['encoder'].items()}
    path = str(tmp_path / 'bad.pth.tar')
    torch.save(bad, path)
    enc, pred, tgt, opt, scaler = build_micro()
    with pytest.raises(RuntimeError):
        load_checkpoint(path, enc, pred, tgt, opt, scaler)
    # an unreadable file keeps the reference's behaviour: logged, epoch 0
    *_, epoch = load_checkpoint(str(tmp_path / 'missing.pth.tar'), enc, pred, tgt, opt, scaler)
    assert epoch == 0


def test_load_pretrained_strips_wrapper_prefixes():
    from avjepa_b200.evals.video_classification_frozen.eval import bare_backbone_keys, load_pretrained
    from avjepa_b200.src.models.audiovision_transformer import AudioVisionTransformer
    ck = torch.load(CK, map_location='cpu')
    torch.manual_seed(3)
    bare = AudioVisionTransformer(mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), uniform_power=True, **MICRO)
    load_pretrained(bare, CK, checkpoint_key='target_encoder')
    want = bare_backbone_keys(ck['target_encoder'])
    assert set(want) == set(bare.state_dict())
    for k, v in bare.state_dict().items():
        assert torch.equal(v, want[k]), k
    load_pretrained(bare, CK, checkpoint_key='no_such_key')                        # falls back to 'encoder'
    assert torch.equal(bare.state_dict()['norm.weight'], ck['encoder']['module.backbone.norm.weight'])


def test_checkpoint_written_here_loads_into_the_live_reference(tmp_path):
    """Reverse direction: save_checkpoint(reference_keys=True) -> the reference's own strict load_checkpoint."""
    ref_root = REFERENCE if os.path.isdir(REFERENCE) else os.path.join(os.path.dirname(GOLDEN), '..', 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'app')):
        pytest.skip('reference not available')
    sys.path.insert(0, ref_root)
    try:
        import logging
        logging.disable(logging.CRITICAL)
        from app.avjepa.utils import init_opt as ref_init_opt, load_checkpoint as ref_load
        from src.models.audiovision_transformer import AudioVisionTransformer as RefEnc
        from src.models.audiovisionpredictor import AudioVisionTransformerPredictor as RefPred
        from src.models.utils.multimask import AudioVideoMultiMaskWrapper as RefW, PredictorMultiMaskWrapper as RefPW
    finally:
        sys.path.remove(ref_root)
    from avjepa_b200.app.avjepa.train import save_checkpoint
    from avjepa_b200.app.avjepa.utils import load_checkpoint
    enc, pred, tgt, opt, scaler = build_micro()
    load_checkpoint(CK, enc, pred, tgt, opt, scaler)

    class _Step(object):
        pass
    st = _Step()
    st.encoder, st.predictor, st.target_encoder, st.optimizer, st.scaler = enc, pred, tgt, opt, scaler
    path = str(tmp_path / 'ours.pth.tar')
    save_checkpoint(path, st, 1, 0.5, 2, 1, 6.25e-4, reference_keys=True)
    ln = partial(nn.LayerNorm, eps=1e-6)
    torch.manual_seed(11)
    renc = RefW(RefEnc(mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, **MICRO))
    rpred = RefPW(RefPred(img_size=32, patch_size=16, num_frames=4, tubelet_size=2, embed_dim=16, predictor_embed_dim=8, depth=1,
                          num_heads=2, mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True, use_mask_tokens=True,
                          num_mask_tokens=2))
    rtgt = copy.deepcopy(renc)
    ropt, rscaler, _, _ = ref_init_opt(encoder=renc, predictor=rpred, wd=0.04, final_wd=0.4, start_lr=2e-4, ref_lr=6.25e-4,
                                       final_lr=1e-6, iterations_per_epoch=3, warmup=1, num_epochs=2, ipe_scale=1.25,
                                       mixed_precision=False)
    renc, rpred, rtgt = (torch.nn.DataParallel(m) for m in (renc, rpred, rtgt))
    *_, epoch = ref_load(path, renc, rpred, rtgt, ropt, rscaler)
    logging.disable(logging.NOTSET)
    assert epoch == 1                                           # the reference swallows failures into epoch 0
    for k, v in renc.state_dict().items():
        assert torch.equal(v, enc.state_dict()[k[len('module.'):]]), k
    # the reference's AdamW steps correctly from our serialised state (independent step counters)
    for g in ropt.param_groups:
        for p in g['params']:
            if p.requires_grad:
                p.grad = torch.full_like(p, 1e-3)
    ropt.step()
    assert all(float(s['step']) == 3.0 for s in ropt.state.values())
