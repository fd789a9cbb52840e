"""Host logic: positional tables, schedules, init parity, state-dict contract, C-ABI surface."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT
from helpers import build_product, checksum, golden


def test_posemb_bit_exact():
    from avjepa_b200.src.models.utils import pos_embs
    from oracle import avjepa_oracle as O
    g = golden('posemb.npz')
    a = pos_embs.get_3d_sincos_pos_embed(1024, 14, 8, uniform_power=True)[::97, ::13].astype(np.float32)
    assert np.array_equal(a, g['v3d_1024_up'])
    assert np.array_equal(pos_embs.get_3d_sincos_pos_embed(192, 14, 8)[::97, ::7].astype(np.float32), g['v3d_192'])
    assert np.array_equal(pos_embs.get_2d_sincos_pos_embed_xy(384, 8, 12)[::5, ::11].astype(np.float32), g['a2d_384'])
    assert np.array_equal(pos_embs.get_2d_sincos_pos_embed(768, 14)[::9, ::17].astype(np.float32), g['i2d_768'])
    assert np.array_equal(O.sincos_3d(1024, 14, 8, True)[::97, ::13].astype(np.float32), g['v3d_1024_up'])
    assert np.array_equal(O.sincos_2d_xy(384, 8, 12)[::5, ::11].astype(np.float32), g['a2d_384'])


def test_schedules_match_oracle_and_reference_formula():
    from avjepa_b200.src.utils.schedulers import CosineWDSchedule, WarmupCosineSchedule
    from oracle import avjepa_oracle as O

    class Opt:
        param_groups = [{'lr': 0, 'weight_decay': 0}, {'lr': 0, 'weight_decay': 0, 'WD_exclude': True}]

    opt = Opt()
    t_max = int(1.25 * 300 * 300)
    s = WarmupCosineSchedule(opt, warmup_steps=40 * 300, start_lr=2e-4, ref_lr=6.25e-4, final_lr=1e-6, T_max=t_max)
    w = CosineWDSchedule(opt, ref_wd=0.04, final_wd=0.4, T_max=t_max)
    for i in range(1, 13000):
        lr, wd = s.step(), w.step()
        if i in (1, 2, 11999, 12000, 12001, 12999):
            assert lr == O.lr_at(i, 40 * 300, 2e-4, 6.25e-4, 1e-6, t_max)
            assert wd == O.wd_at(i, 0.04, 0.4, t_max)
    assert opt.param_groups[0]['lr'] == lr and opt.param_groups[1]['lr'] == lr
    assert opt.param_groups[0]['weight_decay'] == wd and opt.param_groups[1]['weight_decay'] == 0
    assert s.value(1) == pytest.approx(2e-4 + (1 / 12000) * (6.25e-4 - 2e-4))


def test_init_is_bit_identical_to_reference_init():
    """Same seed => same parameters as the reference's init_audio_video_model (checksums from
    the live reference in tests/golden/init_tiny.npz)."""
    enc, pred = build_product('vit_tiny', seed=0)
    g = golden('init_tiny.npz')
    names = set()
    for tag, m in (('enc', enc), ('pred', pred)):
        for n, p in m.named_parameters():
            key = f'{tag}.{n}'
            names.add(key)
            assert key in g.files, f'parameter {key} does not exist in the reference'
            assert np.array_equal(checksum(p), g[key]), key
    assert names == set(g.files)


def test_state_dict_contract():
    enc, pred = build_product('vit_tiny', seed=0)
    sd = enc.state_dict()
    assert sd['backbone.video_pos_embed'].shape == (1, 1568, 192)
    assert sd['backbone.audio_pos_embed'].shape == (1, 96, 192)
    assert sd['backbone.patch_embed.proj.weight'].shape == (192, 3, 2, 16, 16)
    assert sd['backbone.patch_embed.audio_proj.weight'].shape == (192, 1, 16, 16)
    assert sd['backbone.blocks.11.attn.qkv.weight'].shape == (576, 192)
    assert sd['backbone.blocks.0.mlp.fc1.weight'].shape == (768, 192)
    sp = pred.state_dict()
    assert sp['backbone.predictor_embed_v.weight'].shape == (384, 192)
    assert sp['backbone.mask_tokens_a.1'].shape == (1, 1, 384)
    assert sp['backbone.predictor_pos_embed_v'].shape == (1, 1568, 384)
    assert sp['backbone.predictor_proj.weight'].shape == (192, 384)
    assert sum(p.numel() for p in enc.parameters()) == 6002688
    assert sum(p.numel() for p in pred.parameters()) == 22156992
    assert enc.backbone.embed_dim == 192 and enc.backbone.num_heads == 3 and enc.backbone.get_num_layers() == 12


def test_video_only_models_construct():
    import avjepa_b200.src.models.vision_transformer as vit
    from avjepa_b200.src.models.predictor import vit_predictor
    m = vit.vit_tiny(img_size=224, num_frames=16, tubelet_size=2, uniform_power=True, use_sdpa=True)
    assert m.pos_embed.shape == (1, 1568, 192) and not m.pos_embed.requires_grad
    p = vit_predictor(img_size=224, num_frames=16, tubelet_size=2, embed_dim=192, predictor_embed_dim=384, depth=2,
                      num_heads=3, use_mask_tokens=True, num_mask_tokens=2)
    assert set(k for k in p.state_dict() if 'mask_tokens' in k) == {'mask_tokens.0', 'mask_tokens.1'}
    assert vit.VIT_EMBED_DIMS['vit_huge'] == 1280


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference tree not mounted')
def test_state_dict_keys_equal_live_reference():
    import sys
    sys.path.insert(0, REFERENCE)
    import logging
    logging.disable(logging.CRITICAL)
    from app.avjepa.utils import init_audio_video_model as ref_init
    torch.manual_seed(0)
    renc, rpred = ref_init(device=torch.device('cpu'), model_name='vit_tiny', pred_depth=2, pred_embed_dim=384,
                           uniform_power=True, use_mask_tokens=True, num_mask_tokens=2)
    enc, pred = build_product('vit_tiny', seed=0, pred_depth=2)
    for a, b in ((renc, enc), (rpred, pred)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
    b.load_state_dict(a.state_dict())       # reference checkpoint loads into the product module


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'avjepa_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(avj_[a-z0-9_]+)\s*\(', src)))


def test_cabi_exports_every_declared_symbol():
    from avjepa_b200 import _cabi
    syms = _header_symbols()
    assert len(syms) >= 25
    assert os.path.exists(_cabi.LIB_PATH), 'libavjepa_sm100.so is not built (run __graft_entry__.build())'
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/avjepa_b200.h but not exported'
    assert set(_cabi.PROTOTYPES) == set(syms), 'ctypes prototypes and header disagree'
    assert lib.avj_version() == 1


def test_missing_library_fails_loudly(monkeypatch):
    from avjepa_b200 import _cabi
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', '/nonexistent/libavjepa_sm100.so')
    with pytest.raises(_cabi.AvjError, match='no CPU fallback'):
        _cabi.load()


def test_cpu_tensors_are_rejected():
    """No silent CPU path: product ops refuse CPU tensors."""
    from avjepa_b200 import _cabi
    from avjepa_b200.src.masks.utils import apply_masks
    with pytest.raises(_cabi.AvjError):
        apply_masks(torch.zeros(1, 4, 8), [torch.zeros(1, 2, dtype=torch.int64)])


def test_video_only_init_is_bit_identical_to_reference_init():
    """init_video_model (app/vjepa/utils.py:86-153): same seed => same 182 parameter tensors as the reference's
    video-only factory (checksums from the live reference in tests/golden/init_video_tiny.npz)."""
    import logging
    logging.disable(logging.CRITICAL)
    from avjepa_b200.app.vjepa.utils import init_video_model
    torch.manual_seed(0)
    np.random.seed(0)
    enc, pred = init_video_model(
        device=torch.device('cpu'), patch_size=16, num_frames=16, tubelet_size=2, model_name='vit_tiny', crop_size=224,
        pred_depth=6, pred_embed_dim=384, uniform_power=True, use_mask_tokens=True, num_mask_tokens=2,
        zero_init_mask_tokens=True, use_sdpa=True)
    g = golden('init_video_tiny.npz')
    names = set()
    for tag, m in (('enc', enc), ('pred', pred)):
        for n, p in m.named_parameters():
            key = f'{tag}.{n}'
            names.add(key)
            assert key in g.files, f'parameter {key} does not exist in the reference'
            assert np.array_equal(checksum(p), g[key]), key
    assert names == set(g.files)
    assert enc.backbone.pos_embed.shape == (1, 1568, 192) and pred.backbone.predictor_pos_embed.shape == (1, 1568, 384)
