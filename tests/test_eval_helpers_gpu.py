"""Frozen-eval consumers and measurement hooks on the GPU: rebuild_tokens (app/avprediction/utils.py:206-231)
against the oracle restatement, the per-launch timing dump of the C ABI, and the TrainStep host read."""
import csv
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _complementary_masks(g, B, n_v=1568, n_a=96, k_v=400, k_a=20):
    def split(n, k):
        keep, rest = [], []
        for _ in range(B):
            p = torch.randperm(n, generator=g)
            keep.append(p[:k].sort().values)
            rest.append(p[k:].sort().values)
        return torch.stack(keep), torch.stack(rest)
    ev, pv = split(n_v, k_v)
    ea, pa = split(n_a, k_a)
    return (ev, ea), (pv, pa)


def test_rebuild_tokens_bit_exact():
    from avjepa_b200.app.avprediction.utils import rebuild_tokens
    from oracle import avjepa_oracle as orc
    g = torch.Generator().manual_seed(3)
    B, D = 3, 192
    ctxt, pred, m_enc, m_pred = [], [], [], []
    for kv, ka in ((400, 20), (111, 48)):
        me, mp = _complementary_masks(g, B, k_v=kv, k_a=ka)
        m_enc.append(me)
        m_pred.append(mp)
        ctxt.append(torch.randn(B, kv + ka, D, generator=g))
        pred.append(torch.randn(B, 1664 - kv - ka, D, generator=g).bfloat16())
    ref = orc.rebuild_tokens(ctxt, pred, m_enc, m_pred)
    dev = torch.device('cuda')
    out = rebuild_tokens([c.to(dev) for c in ctxt], [p.to(dev) for p in pred],
                         [tuple(m.to(dev) for m in me) for me in m_enc], [tuple(m.to(dev) for m in mp) for mp in m_pred])
    for o, r in zip(out, ref):
        assert o.dtype == torch.float32 and tuple(o.shape) == tuple(r.shape)
        assert torch.equal(o.cpu(), r)
    with pytest.raises(IndexError):
        bad = (m_enc[0][0].to(dev), m_enc[0][1].to(dev) + 200)
        rebuild_tokens([ctxt[0].to(dev)], [pred[0].to(dev)], [bad], [tuple(m.to(dev) for m in m_pred[0])])


def test_prof_dump_lists_every_launch(tmp_path):
    from avjepa_b200 import _cabi, engine
    a = torch.randn(256, 128, device='cuda').bfloat16()
    b = torch.randn(192, 128, device='cuda').bfloat16()
    c = torch.empty(256, 192, device='cuda', dtype=torch.bfloat16)
    x = torch.randn(100, 384, device='cuda')
    y = torch.empty(100, 384, device='cuda', dtype=torch.bfloat16)
    mean, rstd = torch.empty(100, device='cuda'), torch.empty(100, device='cuda')
    gam, bet = torch.ones(384, device='cuda'), torch.zeros(384, device='cuda')
    _cabi.prof_enable(True)
    try:
        for _ in range(3):
            engine.gemm(engine.MODE_BF16, _cabi.GEMM_NT, a.data_ptr(), b.data_ptr(), c.data_ptr(), 256, 192, 128, 128, 128, 192,
                        _cabi.BF16)
        engine.layernorm_fwd(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), y.data_ptr(), _cabi.BF16, mean.data_ptr(),
                             rstd.data_ptr(), 100, 384, 1e-6)
        torch.cuda.synchronize()
        fam = _cabi.prof_collect()
        path = tmp_path / 'prof.csv'
        _cabi.prof_dump(path)
    finally:
        _cabi.prof_enable(False)
    rows = list(csv.DictReader(open(path)))
    gemm = [r for r in rows if r['family'] == '0']
    ln = [r for r in rows if r['family'] == '3']
    assert len(gemm) == 3 and len(ln) == 1 and fam['gemm'][2] == 3 and fam['layernorm_fwd'][2] == 1
    assert all((int(r['d1']), int(r['d2']), int(r['d3'])) == (256, 192, 128) and float(r['ms']) > 0 for r in gemm)
    assert float(gemm[0]['work']) == pytest.approx(2.0 * 256 * 192 * 128, rel=1e-5)
    assert (int(ln[0]['d0']), int(ln[0]['d1'])) == (100, 384)
    assert torch.allclose(c.float(), a.float() @ b.float().t(), rtol=2e-2, atol=2e-1)


def test_trainstep_host_read_matches_device_values():
    """sync=True returns the loss scalars through the staged D2H copy; they must equal the device tensors of the
    same step run with sync=False from the same state."""
    import copy
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import step_support as ss
    dev = torch.device('cuda')
    step_a, batch = ss.build_tiny_step(dev, seed=0)
    step_b, _ = ss.build_tiny_step(dev, seed=0)
    out_sync = step_a(*batch, epoch=0, sync=True)
    out_dev = step_b(*batch, epoch=0, sync=False)
    torch.cuda.synchronize()
    for hs, dv in zip(out_sync[:3], out_dev[:3]):
        assert isinstance(hs, float)
        assert abs(hs - float(dv.detach())) <= 1e-6 * max(1.0, abs(hs))
    assert out_sync[3:] == out_dev[3:]
