"""Frozen-feature paths of the encoder against the UNMODIFIED reference modules (installed in baseline/_ref, present on
the GPU box): ``out_layers`` (per-layer normed features, ``src/models/audiovision_transformer.py:226-236``) and
``interpolate_pos_encoding`` (clips at another resolution, ``:241-270``).  Same weights in both models (state dicts are
key-compatible), reference in fp32 on the CPU, product in fp32 check mode (1e-4) and bf16 (2e-2)."""
import pytest
import torch

from helpers import build_product, rel_err, step_inputs

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _reference_backbone(product_wrapper):
    from baseline import reference_step as R
    if not R.available():
        pytest.skip('baseline/_ref (the installed reference) is not present')
    R._import_reference()
    import src.models.audiovision_transformer as ref_vit          # the reference's module, from baseline/_ref
    assert ref_vit.__file__.startswith(R.REF_DIR)
    ref = ref_vit.vit_tiny(img_size=224, patch_size=16, num_frames=16, tubelet_size=2, uniform_power=True, use_sdpa=True)
    sd = {k[len('backbone.'):]: v.detach().cpu().float() for k, v in product_wrapper.state_dict().items()}
    ref.load_state_dict(sd, strict=True)
    return ref.eval()


def test_out_layers_match_the_reference_module():
    clips, asgram, _, _ = step_inputs()
    enc, _ = build_product('vit_tiny', seed=0, device=DEV)
    ref = _reference_backbone(enc)
    layers = [2, 7, 11]
    ref.out_layers = layers
    enc.backbone.out_layers = layers
    with torch.no_grad():
        want = ref(clips, asgram)
        got = enc.backbone(clips.to(DEV), asgram.to(DEV))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            got16 = enc.backbone(clips.to(DEV), asgram.to(DEV))
    assert len(want) == len(got) == len(got16) == 3
    for w, g, g16 in zip(want, got, got16):
        assert g.shape == w.shape
        assert rel_err(g, w) < 1e-4
        assert rel_err(g16, w) < 2e-2


def test_interpolated_positional_embedding_matches_the_reference_module():
    g = torch.Generator().manual_seed(5)
    clips = torch.randn(2, 3, 16, 160, 192, generator=g)            # 8 x 10 x 12 tokens instead of 8 x 14 x 14
    asgram = -80.0 * torch.rand(2, 1, 128, 192, generator=g)
    enc, _ = build_product('vit_tiny', seed=0, device=DEV)
    ref = _reference_backbone(enc)
    with torch.no_grad():
        want = ref(clips, asgram)
        got = enc.backbone(clips.to(DEV), asgram.to(DEV))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            got16 = enc.backbone(clips.to(DEV), asgram.to(DEV))
    assert want.shape == got.shape == (2, 8 * 10 * 12 + 96, 192)
    assert rel_err(got, want) < 1e-4
    assert rel_err(got16, want) < 2e-2
