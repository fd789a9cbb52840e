"""Step-level parity at the BASELINE.json configurations (not only ViT-tiny): the product path in bf16 -- D = 768 / 1024
GEMM shapes, 12 / 24 layers of bf16 error growth, the head_dim 64 encoder attention AND the head_dim 32 / 24 predictor
attention on the tcgen05 kernels, and head_dim 80 (ViT-H) -- against the fp32 CPU oracle on the same seeded inputs and bit-identical weights.

Bars (north-star): loss within 1e-2 relative of the oracle; global gradient relative error
<= max(1e-2, 1.25 x the error of the oracle's own arithmetic under bf16 autocast on the same GPU)."""
import json

import pytest

import step_support as S

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _check(res):
    print(json.dumps(res))
    assert res['loss_rel'] <= 1e-2, res
    assert res['grad_rel'] <= max(1e-2, 1.25 * res['autocast_grad_rel']), res
    assert res['z_rel'] <= 3e-2 and res['h_rel'] <= 3e-2, res


def test_vit_base_bf16_step_matches_oracle():
    """BASELINE config 2 shape: ViT-B/16 (D 768, 12 layers, 12 heads: hd 64; predictor 384/12 heads: hd 32), depth-12 predictor."""
    _check(S.config_parity('vit_base', 12, DEV))


def test_vit_large_bf16_step_matches_oracle():
    """BASELINE config 3 shape: ViT-L/16 (D 1024, 24 layers, 16 heads: hd 64; predictor 384/16 heads: hd 24)."""
    _check(S.config_parity('vit_large', 16, DEV))


def test_vit_huge_bf16_step_matches_oracle():
    """BASELINE config 4 shape: ViT-H/16 (D 1280, 32 layers, 16 heads: head_dim 80 -- the two-half tcgen05 attention tiles in a
    full forward + backward), depth-12 predictor."""
    _check(S.config_parity('vit_huge', 16, DEV))


def test_vit_small_fp32_check_mode_depth12_predictor():
    """fp32 check mode above tiny: D 384, hd 64 in both stacks, 1e-4 on loss and gradients."""
    res = S.config_parity('vit_small', 6, DEV, mixed=False)
    print(json.dumps(res))
    assert res['loss_rel'] <= 1e-4 and res['grad_rel'] <= 1e-4, res
