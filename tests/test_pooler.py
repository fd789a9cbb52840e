"""Attentive pooler: the oracle restatement is pinned against the live reference module on the
CPU; the product path (few-query cross-attention kernel + GEMM schedule) is checked against the oracle on the GPU,
forward and backward, complete block and bare cross-attention, depth 1 and 2, fp32 check mode and bf16."""
import os
import sys

import pytest
import torch

from conftest import REFERENCE
from helpers import rel_err


def _product_pooler(seed, **kw):
    from avjepa_b200.src.models.attentive_pooler import AttentivePooler
    torch.manual_seed(seed)
    return AttentivePooler(**kw)


def test_pooler_init_and_oracle_match_the_live_reference():
    if not os.path.isdir(REFERENCE):
        pytest.skip('reference not mounted')
    sys.path.insert(0, REFERENCE)
    try:
        from src.models.attentive_pooler import AttentivePooler as Ref
    finally:
        sys.path.remove(REFERENCE)
    from oracle import avjepa_oracle as O
    for kw in (dict(num_queries=1, embed_dim=64, num_heads=4, depth=1), dict(num_queries=2, embed_dim=64, num_heads=2, depth=2),
               dict(num_queries=1, embed_dim=64, num_heads=4, depth=1, complete_block=False)):
        torch.manual_seed(5)
        ref = Ref(**kw)
        ours = _product_pooler(5, **kw)
        sd_r, sd_o = ref.state_dict(), ours.state_dict()
        assert list(sd_r) == list(sd_o)
        for k in sd_r:
            assert torch.equal(sd_r[k], sd_o[k]), k                     # same-seed initialisation is bit-identical
        x = torch.randn(3, 50, 64)
        want = ref(x)
        got = O.attentive_pooler_forward({k: v for k, v in sd_r.items()}, x, kw['num_heads'], kw.get('complete_block', True))
        assert rel_err(got, want) < 1e-5, kw


@pytest.mark.gpu
@pytest.mark.parametrize('kw', [dict(num_queries=1, embed_dim=192, num_heads=3, depth=1), dict(num_queries=2, embed_dim=128, num_heads=2, depth=2),
                                dict(num_queries=1, embed_dim=192, num_heads=12, depth=1, complete_block=False),
                                dict(num_queries=1, embed_dim=640, num_heads=8, depth=1)],
                         ids=['hd64', 'n2_depth2', 'bare_hd16', 'hd80'])
def test_pooler_matches_oracle_on_gpu(kw):
    from oracle import avjepa_oracle as O
    heads, cb = kw['num_heads'], kw.get('complete_block', True)
    D = kw['embed_dim']
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 200, D, generator=g)
    for mixed, tol_f, tol_g in ((False, 1e-4, 1e-3), (True, 2e-2, 4e-2)):
        ours = _product_pooler(7, **kw).cuda()
        # give the zero-initialised biases and the LayerNorm weights some structure
        torch.manual_seed(11)
        for n, p in ours.named_parameters():
            if p.dim() == 1:
                p.data.add_(0.1 * torch.randn_like(p))
        P = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in ours.state_dict().items()}
        xr = x.clone().requires_grad_(True)
        ref = O.attentive_pooler_forward(P, xr, heads, cb)
        (ref * torch.linspace(-1, 1, ref.numel()).reshape(ref.shape)).sum().backward()
        xg = x.cuda().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=mixed):
            out = ours(xg)
        assert out.shape == ref.shape
        assert rel_err(out, ref) < tol_f, (mixed, rel_err(out, ref))
        (out * torch.linspace(-1, 1, ref.numel()).reshape(ref.shape).cuda()).sum().backward()
        assert rel_err(xg.grad, xr.grad) < tol_g, ('dx', mixed, rel_err(xg.grad, xr.grad))
        for n, p in ours.named_parameters():
            want = P[n].grad
            if want is None or float(want.norm()) < 1e-7:
                continue
            assert p.grad is not None, n
            assert rel_err(p.grad, want) < tol_g, (n, mixed, rel_err(p.grad, want))
    # frozen-eval use: no grad, no dx, classifier head on top
    from avjepa_b200.src.models.attentive_pooler import AttentiveClassifier
    torch.manual_seed(1)
    clf = AttentiveClassifier(embed_dim=D, num_heads=heads, depth=1, num_classes=10).cuda()
    with torch.no_grad():
        y = clf(x.cuda())
    assert y.shape == (3, 10) and bool(torch.isfinite(y).all())
